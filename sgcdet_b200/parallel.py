"""Multi-GPU partitioning of the view-transform path.

* **Scene-batch data parallel** (primary): scenes are independent, one (or more) per rank, no data-path collective;
  ``allreduce_gradients`` averages the path's weight gradients with one flattened NCCL all-reduce.
* **View sharding** (config 5): the V views of ONE scene are split over the ranks.  Everything per (view, voxel) --
  feature projection, voxel projection/compaction, lift -- is local to the view's owner.  The cross-view fusion needs
  statistics over all views, so each of its softmax/mean statistics becomes (local partial) -> all-reduce -> (local
  finish): per level 2 exchange steps forward (sum/count, then max followed by the (s, o) partial-softmax sums --
  the log-sum-exp merge) and 2 backward (the softmax-normaliser dot, the query gradient).  The voxel-count
  GEMMs / LN / FFN / upsample / top-k are replicated, so every rank holds the same volume and the same
  (deterministic) selection.

Two drivers of the same kernels:
* ``ViewShardExchange`` + ``AdaptiveSparseHead.forward(..., view_shard=...)``: the product path.  The exchanges are single
  kernel launches over NVLink peer memory inside the fused encoder layer (``functional.EncoderLayerRows``), on the own
  tensor-core GEMMs, side streams and all; the whole step is one CUDA graph.
* ``forward_view_sharded`` + ``Collective``: the reference formulation (eager, library GEMMs, NCCL / gloo or an in-process
  simulation with several shards in one process) the single-GPU and CPU tests check the sharded math with.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from . import functional as SF
from ._lib import call, ptr, stream

F32 = torch.float32
H = SF.NUM_HEADS


class Collective:
    """Reduce a list of per-local-shard tensors to ONE tensor shared by all shards (and all ranks)."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None, use_dist: Optional[bool] = None):
        self.group = group
        self.use_dist = (dist.is_available() and dist.is_initialized()) if use_dist is None else use_dist

    def reduce(self, parts: Sequence[torch.Tensor], op: str) -> torch.Tensor:
        out = parts[0]
        for p in parts[1:]:
            out = torch.maximum(out, p) if op == 'max' else out + p
        if self.use_dist and dist.get_world_size(self.group) > 1:
            out = out.contiguous() if len(parts) == 1 else out
            if len(parts) == 1:
                out = out.clone()
            dist.all_reduce(out, op=dist.ReduceOp.MAX if op == 'max' else dist.ReduceOp.SUM, group=self.group)
        return out


def merge_partial_softmax(m_parts, s_parts, o_parts):
    """Log-sum-exp merge of per-shard partial softmax statistics (reference formulation used by the CPU/gloo tests):
    m = max_g m_g, s = sum_g s_g e^{m_g-m}, o = sum_g o_g e^{m_g-m}; returns (m, s, o).  The CUDA path avoids the
    rescale by exchanging the max first (csrc/sgc_crossview.cu)."""
    m = m_parts[0]
    for p in m_parts[1:]:
        m = torch.maximum(m, p)
    s = sum(sp * torch.exp(mp - m) for mp, sp in zip(m_parts, s_parts))
    o = sum(op * torch.exp(mp - m).unsqueeze(-1) for mp, op in zip(m_parts, o_parts))
    return m, s, o


def allreduce_gradients(params: Sequence[torch.nn.Parameter], group=None, average: bool = True) -> None:
    """Scene-batch DP: one flattened all-reduce over the path's weight gradients (~8 MB at C=256)."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not (dist.is_available() and dist.is_initialized()):
        return
    world = dist.get_world_size(group)
    if world == 1:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    if average:
        flat.div_(world)
    torch._foreach_copy_([g.view(-1) for g in grads], list(flat.split([g.numel() for g in grads])))


def shard_views(num_views: int, world: int, rank: int) -> range:
    """Contiguous, balanced split of the views over the ranks."""
    base, rem = divmod(num_views, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def shard_scene_inputs(mlvl_feats, img_meta, mlvl_dpt_dists, views: range):
    """Slice a full scene down to the views one rank owns (feature maps, depth maps, extrinsics)."""
    idx = slice(views.start, views.stop)
    meta = dict(img_meta)
    l2i = dict(img_meta['lidar2img'])
    l2i['extrinsic'] = list(img_meta['lidar2img']['extrinsic'][idx])
    meta['lidar2img'] = l2i
    meta.pop('sgc_projection', None)
    return [f[:, idx].contiguous() for f in mlvl_feats], meta, [d[:, idx].contiguous() for d in mlvl_dpt_dists]


class CrossViewSharded(torch.autograd.Function):
    """DCA:815-837 with the views split over shards; see the module docstring.  Inputs: per-shard pair lists and
    per-shard slots; output: the fused rows [Q,C] (replicated)."""

    @staticmethod
    def forward(ctx, coll: Collective, pls: List[SF.PairList], w_out, b_out, in_w, in_b, wo, bo, lw, *slots_list):
        Q = pls[0].Q
        C = slots_list[0].shape[1]
        dh = C // H
        scale = 1.0 / math.sqrt(dh)
        dev = slots_list[0].device
        n = len(pls)
        if lw is None:
            lw = SF.LevelWeights(w_out, w_out, in_w, wo, w_out, w_out)
        # exchange 1: sum over views + view count
        sums = []
        for pl, s in zip(pls, slots_list):
            t_ = torch.empty(Q, C, device=dev, dtype=F32)
            call('sgc_crossview_sum_fwd', ptr(s), ptr(pl.pair_index), pl.V, Q, C, ptr(t_), stream())
            sums.append(t_)
        ssum = coll.reduce(sums, 'sum')
        count = coll.reduce([pl.count.to(F32) for pl in pls], 'sum')
        mean = ssum / count.clamp(min=1.0).unsqueeze(1)
        wq, wk, wv = in_w[:C], in_w[C:2 * C] * scale, in_w[2 * C:]
        bq, bv = in_b[:C], in_b[2 * C:]
        g = SF.mm_nt(mean, w_out, lw.w_out) + b_out
        qv = SF.mm_nt(g, wq, lw.wq) + bq
        qt = torch.bmm(SF._heads_cols(qv, 0), lw.wk_rows, out_dtype=F32)
        # exchange 2a: max of the scores; 2b: partial softmax sums (the log-sum-exp merge)
        scores, mloc = [], []
        for pl, s in zip(pls, slots_list):
            sc = torch.empty(pl.cap, H, device=dev, dtype=F32)
            ml = torch.empty(Q, H, device=dev, dtype=F32)
            call('sgc_cvs_scores', ptr(qt), ptr(s), ptr(pl.pair_index), pl.V, Q, C, ptr(sc), ptr(ml), stream())
            scores.append(sc)
            mloc.append(ml)
        m = coll.reduce(mloc, 'max')
        es, sl, ol = [], [], []
        for pl, s, sc in zip(pls, slots_list, scores):
            e = torch.empty(pl.cap, H, device=dev, dtype=F32)
            s_ = torch.empty(Q, H, device=dev, dtype=F32)
            o_ = torch.empty(H, Q, C, device=dev, dtype=F32)
            call('sgc_cvs_accum', ptr(sc), ptr(m), ptr(s), ptr(pl.pair_index), pl.V, Q, C, ptr(e), ptr(s_), ptr(o_), stream())
            es.append(e); sl.append(s_); ol.append(o_)
        ssm = coll.reduce(sl, 'sum')                      # [Q,8]
        osum = coll.reduce(ol, 'sum')                     # [8,Q,C]
        t = (osum / ssm.t().clamp(min=1e-30).unsqueeze(-1)).contiguous()
        o = torch.bmm(SF.split_cols(t.view(H * Q, C), 0).view(H, Q, 3 * C),
                      lw.wv_cols.view(H, dh, 3 * C).transpose(1, 2), out_dtype=F32)
        o2 = o.transpose(0, 1).reshape(Q, C) + bv
        has = (count > 0).to(F32).unsqueeze(1)
        out = (SF.mm_nt(o2, wo, lw.wo) + bo) * has
        ctx.save_for_backward(mean, g, qv, qt, t, ssm, o2, has, count, w_out, in_w, wo, *slots_list, *es)
        ctx.pls, ctx.coll, ctx.lw, ctx.n = pls, coll, lw, n
        return out

    @staticmethod
    def backward(ctx, gout):
        saved = ctx.saved_tensors
        mean, g, qv, qt, t, ssm, o2, has, count, w_out, in_w, wo = saved[:12]
        n = ctx.n
        slots_list, es = saved[12:12 + n], saved[12 + n:12 + 2 * n]
        pls, coll, lw = ctx.pls, ctx.coll, ctx.lw
        Q = pls[0].Q
        C = slots_list[0].shape[1]
        dh = C // H
        scale = 1.0 / math.sqrt(dh)
        dev = gout.device
        wq = in_w[:C]
        gout = gout * has
        g_wo = SF.mm_tn(gout, o2)
        g_bo = SF.colsum(gout)
        go2 = SF.mm_nt(gout, wo.t(), lw.wo_t)
        g_bv = SF.colsum(go2)
        gt = torch.bmm(SF._heads_cols(go2, 0), lw.wv_rows, out_dtype=F32)
        g_wv = torch.bmm(SF._heads_rows_t(go2, 0), SF.split_rows(t.view(H * Q, C), Q, 1), out_dtype=F32).reshape(C, C)
        # exchange 3: the softmax-normaliser dot  D[q,h] = sum_v alpha g_alpha over ALL views
        alphas, galphas, dloc = [], [], []
        for pl, s, e in zip(pls, slots_list, es):
            a = torch.empty(pl.cap, H, device=dev, dtype=F32)
            ga = torch.empty(pl.cap, H, device=dev, dtype=F32)
            d = torch.empty(Q, H, device=dev, dtype=F32)
            call('sgc_cvs_bwd_dot', ptr(s), ptr(e), ptr(ssm), ptr(pl.pair_index), pl.V, Q, C, ptr(gt), ptr(a), ptr(ga), ptr(d),
                 stream())
            alphas.append(a); galphas.append(ga); dloc.append(d)
        dsum = coll.reduce(dloc, 'sum')
        # exchange 4: the query gradient
        gscores, gqts = [], []
        for pl, s, a, ga in zip(pls, slots_list, alphas, galphas):
            gs = torch.empty(pl.cap, H, device=dev, dtype=F32)
            gq = torch.empty(H, Q, C, device=dev, dtype=F32)
            call('sgc_cvs_bwd_qt', ptr(s), ptr(a), ptr(ga), ptr(dsum), ptr(pl.pair_index), pl.V, Q, C, ptr(gs), ptr(gq), stream())
            gscores.append(gs); gqts.append(gq)
        gqt = coll.reduce(gqts, 'sum')
        gqv_h = torch.bmm(SF.split_cols(gqt.view(H * Q, C), 0).view(H, Q, 3 * C),
                          lw.wk_cols.view(H, dh, 3 * C).transpose(1, 2), out_dtype=F32)
        gqv = gqv_h.transpose(0, 1).reshape(Q, C)
        g_wk = torch.bmm(SF._heads_rows_t(qv, 0), SF.split_rows(gqt.view(H * Q, C), Q, 1), out_dtype=F32).reshape(C, C) * scale
        g_wq = SF.mm_tn(gqv, g)
        g_bq = SF.colsum(gqv)
        gg = SF.mm_nt(gqv, wq.t(), lw.wq_t)
        g_wout = SF.mm_tn(gg, mean)
        g_bout = SF.colsum(gg)
        gmean = SF.mm_nt(gg, w_out.t(), lw.w_out_t).contiguous()
        cnt_i = count.to(torch.int32).contiguous()
        gslots = []
        for pl, s, a, gs in zip(pls, slots_list, alphas, gscores):
            gsl = torch.empty_like(s)
            call('sgc_cvs_bwd_slots', ptr(qt), ptr(a), ptr(gs), ptr(pl.pair_index), pl.V, Q, C, ptr(gt), ptr(gmean), ptr(cnt_i),
                 ptr(gsl), stream())
            gslots.append(gsl)
        g_in_w = torch.cat([g_wq, g_wk, g_wv], dim=0)
        g_in_b = torch.cat([g_bq, torch.zeros_like(g_bq), g_bv], dim=0)
        return (None, None, g_wout, g_bout, g_in_w, g_in_b, g_wo, g_bo, None, *gslots)


LIFT_SIDE_PARAMS = ('deformable_attention.value_proj', 'deformable_attention.sampling_offsets',
                    'deformable_attention.sampling_offsets_depth', 'deformable_attention.attention_weights')


def allreduce_view_sharded_gradients(head: torch.nn.Module, group=None) -> None:
    """After a view-sharded backward: the lift-side parameters (value_proj / offset / weight Linear layers) hold
    PARTIAL gradients (sum over the local views) -> all-reduce SUM; every other parameter was computed from
    replicated activations and already holds the full gradient on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    grads = [p.grad for n_, p in head.named_parameters() if p.grad is not None and any(k in n_ for k in LIFT_SIDE_PARAMS)]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    torch._foreach_copy_([g.view(-1) for g in grads], list(flat.split([g.numel() for g in grads])))


class ViewShardExchange:
    """The exchange steps of view sharding over NVLink peer memory (``sgcdet_b200.peer``): one symmetric buffer per rank,
    every exchange ONE kernel launch, so the whole sharded step -- collectives included -- is captured into one CUDA graph.

    Usage (one process per GPU, every rank holds the views ``shard_views(V, world, rank)`` of the SAME scene)::

        xch = ViewShardExchange(head)                                   # once
        vol, valid, occ = head(feats_local, meta_local, dists_local, view_shard=xch)
        loss.backward()
        xch.reduce_gradients(head)        # lift-side + per-head key / value weights hold partial sums over the views

    The per-voxel chain is replicated: run it in eval mode or with identical RNG state on the ranks (FFN dropout)."""

    def __init__(self, head, group=None, device=None, mem=None):
        """``mem``: an existing ``peer.PeerMemory`` to use instead of allocating one (it has to hold ``required_bytes(head)``)."""
        from . import peer
        if mem is not None:
            self.mem, self.device = mem, mem.device
            return
        self.mem = peer.PeerMemory(self.required_bytes(head), group, device)
        self.device = self.mem.device

    @staticmethod
    def required_bytes(head) -> int:
        C = head.embed_dims
        rows = [head.base_heads[0].num_voxels] + [min(k, h.num_voxels) for k, h in zip(head.topk_list, head.base_heads[1:])]
        rows += [h.num_voxels for h in head.base_heads[1 + len(head.topk_list):]]
        n_grad = sum(p.numel() for p in head.parameters())
        return 4 * max(max(rows) * (C + H) + 16, n_grad + 16)

    def view(self, shape, offset_floats: int = 0) -> torch.Tensor:
        return self.mem.view(shape, 4 * offset_floats)

    def reduce(self, n: int, op: str) -> torch.Tensor:
        out = torch.empty(n, device=self.device, dtype=F32)
        return self.mem.all_reduce(n, out, op)

    @staticmethod
    def partial_gradients(head):
        """The gradient tensors that hold PARTIAL sums over this rank's views after a view-sharded backward: the four
        projection layers of the deformable attention (they act on the per-view feature maps) and the key / value rows of
        ``attention_pooling.in_proj_weight`` (their products run over the per-view softmax partials)."""
        out = []
        for n_, p in head.named_parameters():
            if p.grad is None:
                continue
            if any(k in n_ for k in LIFT_SIDE_PARAMS):
                out.append(p.grad.view(-1))
            elif n_.endswith('attention_pooling.in_proj_weight'):
                C = p.shape[1]
                out.append(p.grad[C:].reshape(-1))     # rows [C, 3C): key and value projections (a contiguous view)
        return out

    def reduce_gradients(self, head, sync_replicated: bool = True) -> None:
        """Complete the parameter gradients after a view-sharded backward, ONE all-reduce launch: the partial ones are
        summed over the ranks; with ``sync_replicated`` every other gradient is replaced by rank 0's copy (they are equal
        up to the summation order of the few atomically accumulated ones -- bias and occupancy-head gradients -- and have to
        stay BIT-identical, or the replicated weights and with them the top-k selection drift apart between the ranks)."""
        if self.mem.world == 1:
            return
        partial = {g.data_ptr() for g in self.partial_gradients(head)}
        grads, zero = [], []
        for n_, p in head.named_parameters():
            if p.grad is None:
                continue
            if n_.endswith('attention_pooling.in_proj_weight'):
                C = p.shape[1]
                pieces = [(p.grad[:C].reshape(-1), False), (p.grad[C:].reshape(-1), True)]
            else:
                pieces = [(p.grad.view(-1), p.grad.data_ptr() in partial)]
            for g, is_partial in pieces:
                if is_partial:
                    grads.append(g)
                elif sync_replicated:
                    grads.append(g)
                    if self.mem.rank != 0:
                        zero.append(g)
        if not grads:
            return
        if zero:
            torch._foreach_zero_(zero)         # sum over the ranks == rank 0's values, bit for bit
        sizes = [g.numel() for g in grads]
        n = sum(sizes)
        buf = self.view((n,))
        torch._foreach_copy_(list(buf.split(sizes)), grads)
        red = self.reduce(n, 'sum')
        torch._foreach_copy_(grads, list(red.split(sizes)))

    def close(self):
        self.mem.close()


def forward_view_sharded(head, shards, group=None, forced_selection=None, use_dist: Optional[bool] = None):
    """AdaptiveSparseHead.forward (AdaptiveSparseHead.py:43-93) with the views of the scene split over shards.

    ``shards`` = list of ``(mlvl_feats, img_meta, mlvl_dpt_dists)`` owned by THIS process (one per rank with a real
    process group; several for the in-process simulation).  Returns the same ``(volume, valid, occ_preds)`` on every
    rank."""
    from . import plugin as PL
    coll = Collective(group, use_dist)
    nl = len(head.base_heads)
    meta0 = shards[0][1]
    dev = shards[0][0][0].device
    hws = [(meta0['img_shape'][0] // (4 * 2 ** (nl - 1 - i)), meta0['img_shape'][1] // (4 * 2 ** (nl - 1 - i))) for i in range(nl)]
    projs = [PL.projection_on_device(sh[1], dev) for sh in shards]
    vol = None
    occ_list, masks = [], [None] * nl
    for i in range(nl):
        dh_ = head.base_heads[i]
        fi = nl - 1 - i
        layer = dh_.cross_transformer.encoder.layers[0]
        attn = layer.attentions[0]
        mha = attn.attention_pooling
        dbound = dh_.cross_transformer.encoder.dbound
        sel = None
        if i > 0:
            lin = head.occ_pred_heads[i - 1][0]
            up, occ = SF.UpsampleOcc.apply(vol, lin.weight.view(-1), lin.bias)
            occ_list.append(occ.view(1, -1))
            if forced_selection is not None and forced_selection[i] is not None:
                sel = forced_selection[i]
                mask = torch.zeros(occ.numel(), device=dev, dtype=torch.uint8)
                mask[sel.long()] = 1
            else:
                sel, mask = SF.topk_select(occ, min(head.topk_list[i - 1], occ.numel()))
            masks[i] = mask
        pls, slots_list, lw = [], [], None
        for (feats, meta, dists), proj in zip(shards, projs):
            # the sharded cross-view block runs its voxel-count GEMMs on the library path, which needs the bf16x3 images
            pre = dh_.prepare(feats[fi], dists[fi], hws[i], images=True)
            lw = pre['lw']
            pl = SF.project_compact(proj, dh_.ref_3d, sel, meta, dbound)
            slots, _ = SF.Lift.apply(pre['vg'], pre['dist'], pre['vbias'], pre['gbias'], pl, hws[i][0], hws[i][1])
            pls.append(pl)
            slots_list.append(slots)
        x = CrossViewSharded.apply(coll, pls, attn.output_proj.weight, attn.output_proj.bias, mha.in_proj_weight,
                                   mha.in_proj_bias, mha.out_proj.weight, mha.out_proj.bias, lw, *slots_list)
        x = attn.dropout(x)
        x = layer.norms[0](x)
        x = layer.ffns[0](x, lw=lw)
        x = layer.norms[1](x)
        if i == 0:
            X, Y, Z = (int(v) for v in dh_.n_voxels)
            vol = x.view(X, Y, Z, head.embed_dims)
        else:
            vol = SF.ScatterAddRows.apply(up, x, sel)
    volume_out = vol.permute(3, 0, 1, 2).unsqueeze(0)
    occ_preds = torch.cat(occ_list[::-1], dim=1)
    valid = head.get_valid(masks[nl - 1]).unsqueeze(0).unsqueeze(0).detach()
    return volume_out, valid, occ_preds
