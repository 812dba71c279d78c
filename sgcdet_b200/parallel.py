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

``ViewShardExchange`` + ``AdaptiveSparseHead.forward(..., view_shard=...)`` is the product path: the exchanges are single
kernel launches over NVLink peer memory inside the fused encoder layer (``functional.EncoderLayerRows``), on the own
tensor-core GEMMs, side streams and all; the whole step is one CUDA graph.  ``Collective`` / ``merge_partial_softmax`` are the
host-side reference formulation of the same exchange steps over a torch.distributed group (NCCL or gloo), which the CPU
tests check the merge rules with (tests/test_parallel_cpu.py).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from . import functional as SF

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


LIFT_SIDE_PARAMS = ('deformable_attention.value_proj', 'deformable_attention.sampling_offsets',
                    'deformable_attention.sampling_offsets_depth', 'deformable_attention.attention_weights')


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
