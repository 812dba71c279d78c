"""Host-side orchestration of one view-transform level: autograd Functions around the C-ABI kernels.

PyTorch is plumbing here (device memory, streams, the dense library GEMMs of voxel count / pixel count);
every gather / scatter / projection / selection runs in the hand-written kernels of ``csrc/``.
No function in this module has a CPU or eager fallback.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr, stream

NUM_HEADS = 8
NUM_POINTS = 4
G_CH = 4 * NUM_HEADS * NUM_POINTS  # channels of the folded offset/weight map


# ----------------------------------------------------------------------------------------------
# projection + pair list
# ----------------------------------------------------------------------------------------------

def compute_projection(img_meta: dict, stride: int = 1) -> torch.Tensor:
    """Host-side K[:3,:3]/ratio @ E[:3] in fp32 with the same torch CPU ops as the reference
    (``transformer_utils/encoder.py:168-177``).  Returns [V,3,4] on the CPU."""
    intrinsic = torch.tensor(np.asarray(img_meta['lidar2img']['intrinsic'])[:3, :3])
    ratio = img_meta['ori_shape'][0] / (img_meta['img_shape'][0] / stride)
    intrinsic[:2] /= ratio
    proj = [intrinsic @ torch.tensor(np.asarray(e))[:3] for e in img_meta['lidar2img']['extrinsic']]
    return torch.stack(proj).float()


@dataclass
class PairList:
    """Visible (view, voxel) pairs of one level, view-major (device tensors, no host sync)."""
    V: int
    Q: int
    cap: int
    ref_cam: torch.Tensor       # [V,Q,3] fp32 (u,v,d)
    mask: torch.Tensor          # [V,Q] uint8
    pair_index: torch.Tensor    # [V,Q] int32
    pair_vq: torch.Tensor       # [cap] int32
    view_offsets: torch.Tensor  # [V+1] int32
    count: torch.Tensor         # [Q] int32

    @property
    def n_pairs(self) -> torch.Tensor:
        return self.view_offsets[self.V:]  # 1-element view (device)


def project_compact(proj: torch.Tensor, ref3d: torch.Tensor, sel: Optional[torch.Tensor], img_meta: dict,
                    dbound: Sequence[float], cap: Optional[int] = None) -> PairList:
    """``VoxFormerEncoder_DFA3D.point_sampling`` (encoder.py:179-223) + per-view compaction (DCA:758-773)."""
    dev = ref3d.device
    V = proj.shape[0]
    Q = int(sel.numel()) if sel is not None else int(ref3d.shape[0])
    cap = V * Q if cap is None else cap
    origin = np.asarray(img_meta['lidar2img']['origin'], dtype=np.float32)
    ogfH, ogfW = img_meta['img_shape'][0], img_meta['img_shape'][1]
    eps = 1e-5
    pl = PairList(
        V, Q, cap,
        torch.empty(V, Q, 3, device=dev, dtype=torch.float32),
        torch.empty(V, Q, device=dev, dtype=torch.uint8),
        torch.empty(V, Q, device=dev, dtype=torch.int32),
        torch.empty(cap, device=dev, dtype=torch.int32),
        torch.empty(V + 1, device=dev, dtype=torch.int32),
        torch.empty(Q, device=dev, dtype=torch.int32))
    scratch = torch.empty(_lib.load().sgc_project_scratch_ints(V, Q), device=dev, dtype=torch.int32)
    call('sgc_project_compact', ptr(proj), ptr(ref3d), ptr(sel), V, Q,
         float(origin[0]), float(origin[1]), float(origin[2]),
         float(np.float32(eps)), float(np.float32(1.0 - eps)), float(ogfW), float(ogfH),
         float(np.float32(dbound[0])), float(np.float32(dbound[1] - dbound[0])),
         ptr(pl.ref_cam), ptr(pl.mask), ptr(pl.pair_index), ptr(pl.pair_vq), ptr(pl.view_offsets),
         ptr(pl.count), ptr(scratch), stream())
    return pl


# ----------------------------------------------------------------------------------------------
# dense projection of the feature maps (tensor cores via bf16 hi/lo operand split)
# ----------------------------------------------------------------------------------------------

def _split_weight(w: torch.Tensor):
    hi = w.to(torch.bfloat16)
    lo = (w - hi.float()).to(torch.bfloat16)
    return hi, lo


class ProjectFeatures(torch.autograd.Function):
    """VG[v,s,:] = Wcat @ feat[v,:,s]   with feat the NCHW map cropped to (h,w).

    Wcat [C+128, C] = [value_proj.weight ; folded offset/depth-offset/attention-weight rows] (no bias: the
    biases are applied inside the lift kernel).  Reads NCHW directly (transposed GEMM operand) and writes
    channel-last, so the NCHW->NHWC copy of ``transformer.py:151-170`` never happens.

    Tensor cores at fp32-level accuracy: both operands are split into bf16 hi/lo (csrc/sgc_gemm_prep.cu) and
    hi*hi + lo*hi + hi*lo is evaluated as ONE bf16 GEMM with K concatenated 3x and fp32 accumulation/output
    (library GEMM for now; relative error ~1e-5).
    """

    @staticmethod
    def forward(ctx, feat: torch.Tensor, h: int, w: int, wcat: torch.Tensor):
        V, C, H0, W0 = feat.shape
        S = h * w
        src = feat
        if w != W0 or not feat.is_contiguous():
            src = feat[:, :, :h, :w].contiguous()
            stride = S
        else:
            stride = H0 * W0  # row crop only: the first h*w elements of every channel plane
        acat = torch.empty(V, 3, C, S, device=feat.device, dtype=torch.bfloat16)
        call('sgc_split_bf16x3', ptr(src), V * C, S, stride, C, ptr(acat), stream())
        w_hi, w_lo = _split_weight(wcat)
        bcat = torch.cat([w_hi, w_hi, w_lo], dim=1)  # [N, 3C] pairs with [a_hi | a_lo | a_hi]
        vg = torch.bmm(acat.view(V, 3 * C, S).transpose(1, 2), bcat.t().unsqueeze(0).expand(V, -1, -1),
                       out_dtype=torch.float32)  # [V,S,N]
        ctx.save_for_backward(acat, w_hi, w_lo)
        ctx.dims = (V, C, H0, W0, h, w)
        return vg

    @staticmethod
    def backward(ctx, gvg: torch.Tensor):
        acat, w_hi, w_lo = ctx.saved_tensors
        V, C, H0, W0, h, w = ctx.dims
        S = h * w
        N = w_hi.shape[0]
        gvg = gvg.contiguous()
        gcat = torch.empty(V, S, 3, N, device=gvg.device, dtype=torch.bfloat16)
        call('sgc_split_bf16x3', ptr(gvg), V * S, N, N, 1, ptr(gcat), stream())
        gfeat = gw = None
        if ctx.needs_input_grad[0]:
            wk = torch.cat([w_hi.t(), w_hi.t(), w_lo.t()], dim=1)  # [C, 3N] pairs with [g_hi | g_lo | g_hi]
            g = torch.bmm(wk.unsqueeze(0).expand(V, -1, -1), gcat.view(V, S, 3 * N).transpose(1, 2),
                          out_dtype=torch.float32)  # [V,C,S]
            if h == H0 and w == W0:
                gfeat = g.view(V, C, H0, W0)
            else:
                gfeat = g.new_zeros(V, C, H0, W0)
                gfeat[:, :, :h, :w] = g.view(V, C, h, w)
        if ctx.needs_input_grad[3]:
            g_hi_t = gcat[:, :, 0].transpose(1, 2)  # [V,N,S]
            g_lo_t = gcat[:, :, 1].transpose(1, 2)
            x = torch.bmm(g_hi_t, acat[:, :2].reshape(V, 2 * C, S).transpose(1, 2), out_dtype=torch.float32)  # [V,N,2C]
            y = torch.bmm(g_lo_t, acat[:, 0].transpose(1, 2), out_dtype=torch.float32)  # [V,N,C]
            gw = (x[..., :C] + x[..., C:] + y).sum(0)
        return gfeat, None, None, gw


# ----------------------------------------------------------------------------------------------
# lift
# ----------------------------------------------------------------------------------------------

class Lift(torch.autograd.Function):
    """sgc_lift_fwd / sgc_lift_bwd (see csrc/sgc_lift.cu)."""

    @staticmethod
    def forward(ctx, vg, dist, vbias, gbias, pl: PairList, H: int, W: int):
        V, S, ld = vg.shape
        C = ld - G_CH
        D = dist.shape[-1]
        slots = torch.empty(pl.cap, C, device=vg.device, dtype=torch.float32)
        samp = torch.empty(pl.cap, 32, 4, device=vg.device, dtype=torch.float32)
        base = ptr(vg)
        call('sgc_lift_fwd', base, ld, base + 4 * C, ld, ptr(dist), ptr(vbias), ptr(gbias), ptr(pl.pair_vq),
             ptr(pl.n_pairs), pl.cap, ptr(pl.ref_cam), S, H, W, D, pl.Q, C, ptr(samp), ptr(slots), stream())
        ctx.save_for_backward(vg, dist, vbias, samp)
        ctx.pl, ctx.dims = pl, (S, H, W, D, C)
        ctx.mark_non_differentiable(samp)
        return slots, samp

    @staticmethod
    def backward(ctx, gslots, _gsamp):
        vg, dist, vbias, samp = ctx.saved_tensors
        pl = ctx.pl
        S, H, W, D, C = ctx.dims
        ld = vg.shape[-1]
        gvg = torch.zeros_like(vg)
        gdist = torch.zeros_like(dist)
        gvb = torch.zeros_like(vbias)
        ggb = torch.zeros(G_CH, device=vg.device, dtype=torch.float32)
        base, gbase = ptr(vg), ptr(gvg)
        scratch = torch.empty(_lib.load().sgc_lift_bwd_scratch_floats(pl.cap, C), device=vg.device, dtype=torch.float32)
        call('sgc_lift_bwd', base, ld, base + 4 * C, ld, ptr(dist), ptr(vbias), ptr(pl.pair_vq), ptr(pl.n_pairs),
             pl.cap, ptr(pl.ref_cam), ptr(samp), ptr(gslots.contiguous()), S, H, W, D, pl.Q, C,
             gbase, gbase + 4 * C, ptr(gdist), ptr(gvb), ptr(ggb), ptr(scratch), stream())
        return gvg, gdist, gvb, ggb, None, None, None


# ----------------------------------------------------------------------------------------------
# cross-view fusion
# ----------------------------------------------------------------------------------------------

class CrossView(torch.autograd.Function):
    """DCA:815-837: masked mean over views -> output_proj -> 8-head attention pooling over views.

    The dense projections are voxel-count GEMMs (torch.mm/bmm = library GEMMs) around the two
    cross-view kernels; the backward is written out by hand so that no pair-capacity-sized tensor is ever
    touched outside the kernels.
    """

    @staticmethod
    def forward(ctx, slots, pl: PairList, w_out, b_out, in_w, in_b, wo, bo):
        Q, V = pl.Q, pl.V
        C = slots.shape[1]
        dh = C // NUM_HEADS
        scale = 1.0 / math.sqrt(dh)
        dev = slots.device
        mean = torch.empty(Q, C, device=dev, dtype=torch.float32)
        call('sgc_crossview_mean_fwd', ptr(slots), ptr(pl.pair_index), V, Q, C, ptr(mean), stream())
        wq, wk, wv = in_w[:C], in_w[C:2 * C], in_w[2 * C:]
        bq, bv = in_b[:C], in_b[2 * C:]
        g = torch.addmm(b_out, mean, w_out.t())
        qv = torch.addmm(bq, g, wq.t())
        qv_h = qv.view(Q, NUM_HEADS, dh).transpose(0, 1)  # [8,Q,dh]
        qt = torch.bmm(qv_h, wk.view(NUM_HEADS, dh, C)) * scale  # [8,Q,C]
        t = torch.empty(NUM_HEADS, Q, C, device=dev, dtype=torch.float32)
        alpha = torch.empty(pl.cap, NUM_HEADS, device=dev, dtype=torch.float32)
        call('sgc_crossview_attn_fwd', ptr(qt), ptr(slots), ptr(pl.pair_index), V, Q, C, ptr(t), ptr(alpha), stream())
        o = torch.bmm(t, wv.view(NUM_HEADS, dh, C).transpose(1, 2))  # [8,Q,dh]
        o2 = o.transpose(0, 1).reshape(Q, C) + bv
        has = (pl.count > 0).to(torch.float32).unsqueeze(1)
        out = torch.addmm(bo, o2, wo.t()) * has
        ctx.save_for_backward(slots, mean, g, qv, qt, t, alpha, o2, has, w_out, in_w, wo)
        ctx.pl = pl
        return out

    @staticmethod
    def backward(ctx, gout):
        slots, mean, g, qv, qt, t, alpha, o2, has, w_out, in_w, wo = ctx.saved_tensors
        pl = ctx.pl
        Q, V = pl.Q, pl.V
        C = slots.shape[1]
        dh = C // NUM_HEADS
        scale = 1.0 / math.sqrt(dh)
        dev = slots.device
        wq, wk, wv = in_w[:C], in_w[C:2 * C], in_w[2 * C:]
        gout = gout * has
        g_wo = gout.t() @ o2
        g_bo = gout.sum(0)
        go2 = gout @ wo
        g_bv = go2.sum(0)
        go_h = go2.view(Q, NUM_HEADS, dh).transpose(0, 1)  # [8,Q,dh]
        gt = torch.bmm(go_h, wv.view(NUM_HEADS, dh, C)).contiguous()  # [8,Q,C]
        g_wv = torch.bmm(go_h.transpose(1, 2), t).reshape(C, C)
        gscore = torch.empty(pl.cap, NUM_HEADS, device=dev, dtype=torch.float32)
        gqt = torch.empty(NUM_HEADS, Q, C, device=dev, dtype=torch.float32)
        call('sgc_crossview_attn_bwd_qt', ptr(slots), ptr(alpha), ptr(pl.pair_index), V, Q, C, ptr(gt), ptr(gscore),
             ptr(gqt), stream())
        gqv_h = torch.bmm(gqt, wk.view(NUM_HEADS, dh, C).transpose(1, 2)) * scale  # [8,Q,dh]
        qv_h = qv.view(Q, NUM_HEADS, dh).transpose(0, 1)
        g_wk = (torch.bmm(qv_h.transpose(1, 2), gqt) * scale).reshape(C, C)
        gqv = gqv_h.transpose(0, 1).reshape(Q, C)
        g_wq = gqv.t() @ g
        g_bq = gqv.sum(0)
        gg = gqv @ wq
        g_wout = gg.t() @ mean
        g_bout = gg.sum(0)
        gmean = (gg @ w_out).contiguous()
        gslots = torch.empty_like(slots)
        call('sgc_crossview_attn_bwd_slots', ptr(qt), ptr(alpha), ptr(gscore), ptr(pl.pair_index), V, Q, C, ptr(gt),
             ptr(gmean), ptr(gslots), stream())
        g_in_w = torch.cat([g_wq, g_wk, g_wv], dim=0)
        g_in_b = torch.cat([g_bq, torch.zeros_like(g_bq), g_bv], dim=0)
        return gslots, None, g_wout, g_bout, g_in_w, g_in_b, g_wo, g_bo


# ----------------------------------------------------------------------------------------------
# sparse volume construction
# ----------------------------------------------------------------------------------------------

class UpsampleOcc(torch.autograd.Function):
    """(up [2X,2Y,2Z,C], occ [8XYZ]) = trilinear x2 of a channel-last volume + Linear(C,1)+Sigmoid
    (AdaptiveSparseHead.py:64-71)."""

    @staticmethod
    def forward(ctx, vol, w_occ, b_occ):
        X, Y, Z, C = vol.shape
        up = torch.empty(2 * X, 2 * Y, 2 * Z, C, device=vol.device, dtype=torch.float32)
        occ = torch.empty(8 * X * Y * Z, device=vol.device, dtype=torch.float32)
        call('sgc_upsample2x_occ_fwd', ptr(vol), X, Y, Z, C, ptr(w_occ), ptr(b_occ), ptr(up), ptr(occ), stream())
        ctx.save_for_backward(vol, w_occ, occ)
        return up, occ

    @staticmethod
    def backward(ctx, gup, gocc):
        vol, w_occ, occ = ctx.saved_tensors
        X, Y, Z, C = vol.shape
        dev = vol.device
        gup = gup.contiguous() if gup is not None else torch.zeros(2 * X, 2 * Y, 2 * Z, C, device=dev)
        gin = torch.empty_like(vol)
        gw = torch.zeros(C, device=dev, dtype=torch.float32)
        gb = torch.zeros(1, device=dev, dtype=torch.float32)
        gpre = torch.empty_like(occ) if gocc is not None else None
        call('sgc_upsample2x_occ_bwd', ptr(vol), X, Y, Z, C, ptr(w_occ), ptr(occ), ptr(gup),
             ptr(gocc.contiguous()) if gocc is not None else None, ptr(gpre), ptr(gin), ptr(gw), ptr(gb), stream())
        return gin, gw.view_as(w_occ), gb


def topk_select(occ: torch.Tensor, k: int):
    """``topk_wo_grad`` (AdaptiveSparseHead.py:9-13) + ``nonzero`` (DenseHead.py:66).  Returns
    (sel [k] int32 ascending, mask [N] uint8).  Deterministic: ties -> lower index."""
    N = occ.numel()
    sel = torch.empty(k, device=occ.device, dtype=torch.int32)
    mask = torch.empty(N, device=occ.device, dtype=torch.uint8)
    call('sgc_topk_select', ptr(occ.detach()), N, k, ptr(sel), ptr(mask), stream())
    return sel, mask


class ScatterAddRows(torch.autograd.Function):
    """vol[sel] += y in place (DenseHead.py:80-81 + AdaptiveSparseHead.py:77)."""

    @staticmethod
    def forward(ctx, vol, y, sel):
        C = vol.shape[-1]
        call('sgc_scatter_add_rows', ptr(vol), ptr(sel), ptr(y.contiguous()), sel.numel(), C, stream())
        ctx.mark_dirty(vol)
        ctx.save_for_backward(sel)
        ctx.C = C
        return vol

    @staticmethod
    def backward(ctx, gvol):
        (sel,) = ctx.saved_tensors
        gvol = gvol.contiguous()
        gy = torch.empty(sel.numel(), ctx.C, device=gvol.device, dtype=torch.float32)
        call('sgc_gather_rows', ptr(gvol), ptr(sel), ptr(gy), sel.numel(), ctx.C, stream())
        return gvol, gy, None
