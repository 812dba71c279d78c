"""CPU restatement of the DFA3D operator (test infrastructure only; see oracle/__init__.py).

Follows, in gather form and vectorised with torch CPU ops (works in fp32 and fp64):

  * depth-score sampling  – ``csrc/common/cuda/ms_depth_score_sample_cuda_kernel.cuh:24-148``
    (corner order TL, TR, BR, BL at ``:89-92``; pixel coords ``loc*size-0.5`` at ``:133-135``;
    whole-sample range test at ``:137``; zero outside the d range at ``:53-87``)
  * depth-weighted multi-scale deformable attention – ``wms_deform_attn_cuda_kernel.cuh:24-80,240-303``
    (2-D range test at ``:289``; corner/depth-score pairing ds[0],ds[1],ds[3],ds[2] at ``:51,58,65,72``)
  * backward – the reference kernels (``wms_deform_attn_cuda_kernel.cuh:82-159,305-531``,
    ``ms_depth_score_sample_cuda_kernel.cuh:150-327``) compute the a.e. analytic gradient of the
    forward; here it is obtained by autograd through the restated forward, and stitched the way
    ``multi_scale_3ddeformable_attn_function.py:303-351`` does (uv-grads added onto the 3-D loc grad).
  * an independent second opinion: DFA3D == sum_p w_p * grid_sample(value (x) dist).

Layouts (all contiguous):
  value [B,S,M,Cm]   dist [B,S,M,D]   shapes3d [L,3] int64 (H,W,D)   lsi [L] int64
  loc [B,Q,M,L,P,3] in [0,1] with (x=w, y=h, z=d)   attn [B,Q,M,L,P]   depth_score [B,Q,M,L,P,4]
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _pix(loc_comp: torch.Tensor, size: int) -> torch.Tensor:
    # DSK:133-135 / WMSK:286-287:  x_im = loc * size - 0.5  (product rounded first, then the subtract)
    return loc_comp * float(size) - 0.5


def depth_score_sample_forward(dist, shapes3d, lsi, loc):
    """``ms_depth_score_sample_forward`` (pybind.cpp:31-33) -> [B,Q,M,L,P,4]."""
    B, S, M, Dch = dist.shape
    _, Q, _, L, P, _ = loc.shape
    b_idx = torch.arange(B).view(B, 1, 1, 1)
    m_idx = torch.arange(M).view(1, 1, M, 1)
    per_level = []
    for l in range(L):
        H, W, Dl = (int(x) for x in shapes3d[l])
        start = int(lsi[l])
        w = _pix(loc[:, :, :, l, :, 0], W)
        h = _pix(loc[:, :, :, l, :, 1], H)
        d = _pix(loc[:, :, :, l, :, 2], Dl)
        inr = (h > -1) & (w > -1) & (d > -1) & (h < H) & (w < W) & (d < Dl)  # DSK:137
        h0 = torch.floor(h)
        w0 = torch.floor(w)
        d0 = torch.floor(d)
        ld = d - d0
        hd = 1 - ld
        h0i, w0i, d0i = h0.long(), w0.long(), d0.long()
        d_lo_ok = d0i >= 0
        d_hi_ok = (d0i + 1) <= (Dl - 1)
        d_lo_c = d0i.clamp(0, Dch - 1)
        d_hi_c = (d0i + 1).clamp(0, Dch - 1)
        scores = []
        # corner order of the output: TL, TR, BR, BL  (DSK:89-92)
        for (dh, dw) in ((0, 0), (0, 1), (1, 1), (1, 0)):
            hc = h0i + dh
            wc = w0i + dw
            ok = (hc >= 0) & (hc <= H - 1) & (wc >= 0) & (wc <= W - 1) & inr
            pos = start + hc.clamp(0, H - 1) * W + wc.clamp(0, W - 1)
            v_lo = dist[b_idx, pos, m_idx, d_lo_c] * (ok & d_lo_ok).to(dist.dtype)
            v_hi = dist[b_idx, pos, m_idx, d_hi_c] * (ok & d_hi_ok).to(dist.dtype)
            scores.append(v_lo * hd + v_hi * ld)
        per_level.append(torch.stack(scores, dim=-1))  # [B,Q,M,P,4]
    return torch.stack(per_level, dim=3)  # [B,Q,M,L,P,4]


def wms_deform_attn_forward(value, shapes2d, lsi, loc2d, attn, depth_score):
    """``wms_deform_attn_forward`` (pybind.cpp:20-23) -> [B,Q,M*Cm]."""
    B, S, M, Cm = value.shape
    _, Q, _, L, P, _ = loc2d.shape
    b_idx = torch.arange(B).view(B, 1, 1, 1)
    m_idx = torch.arange(M).view(1, 1, M, 1)
    out = value.new_zeros(B, Q, M, Cm)
    for l in range(L):
        H, W = int(shapes2d[l][0]), int(shapes2d[l][1])
        start = int(lsi[l])
        w = _pix(loc2d[:, :, :, l, :, 0], W)
        h = _pix(loc2d[:, :, :, l, :, 1], H)
        inr = (h > -1) & (w > -1) & (h < H) & (w < W)  # WMSK:289
        h0 = torch.floor(h)
        w0 = torch.floor(w)
        lh = h - h0
        lw = w - w0
        hh = 1 - lh
        hw = 1 - lw
        h0i, w0i = h0.long(), w0.long()
        ds = depth_score[:, :, :, l]  # [B,Q,M,P,4]
        # (dh, dw, bilinear weight, depth-score slot)  -- WMSK:47-76
        taps = ((0, 0, hh * hw, 0), (0, 1, hh * lw, 1), (1, 0, lh * hw, 3), (1, 1, lh * lw, 2))
        acc = 0
        for dh, dw, bw, slot in taps:
            hc = h0i + dh
            wc = w0i + dw
            ok = (hc >= 0) & (hc <= H - 1) & (wc >= 0) & (wc <= W - 1) & inr
            pos = start + hc.clamp(0, H - 1) * W + wc.clamp(0, W - 1)
            v = value[b_idx, pos, m_idx]  # [B,Q,M,P,Cm]
            wgt = bw * ds[..., slot] * ok.to(value.dtype)
            acc = acc + wgt.unsqueeze(-1) * v
        out = out + (acc * attn[:, :, :, l].unsqueeze(-1)).sum(dim=3)
    return out.reshape(B, Q, M * Cm)


def dfa3d_forward(value, dist, shapes3d, lsi, loc, attn):
    """One-stage operator: ``MultiScale3DDeformableAttnFunction_fp32.forward`` (F3D:277-302)."""
    ds = depth_score_sample_forward(dist, shapes3d, lsi, loc)
    out = wms_deform_attn_forward(value, shapes3d[:, :2], lsi, loc[..., :2], attn, ds)
    return out, ds


def wms_deform_attn_backward(value, shapes2d, lsi, loc2d, attn, depth_score, grad_output):
    """Returns (grad_value, grad_sampling_loc[...,2], grad_attn_weight, grad_depth_score) as
    ``wms_deform_attn_backward`` accumulates them (pybind.cpp:25-30)."""
    ins = [t.detach().clone().requires_grad_(True) for t in (value, loc2d, attn, depth_score)]
    out = wms_deform_attn_forward(ins[0], shapes2d, lsi, ins[1], ins[2], ins[3])
    return torch.autograd.grad(out, ins, grad_output.reshape(out.shape))


def depth_score_sample_backward(dist, shapes3d, lsi, loc, grad_depth_score):
    """Returns (grad_dist, grad_sampling_loc[...,3]); the w,h components are identically zero
    (DSK:238-240)."""
    ins = [t.detach().clone().requires_grad_(True) for t in (dist, loc)]
    out = depth_score_sample_forward(ins[0], shapes3d, lsi, ins[1])
    return torch.autograd.grad(out, ins, grad_depth_score)


def dfa3d_backward(value, dist, shapes3d, lsi, loc, attn, grad_output):
    """``MultiScale3DDeformableAttnFunction_fp32.backward`` (F3D:303-351):
    returns (grad_value, grad_dist, grad_loc[...,3], grad_attn)."""
    ds = depth_score_sample_forward(dist, shapes3d, lsi, loc)
    g_value, g_loc2, g_attn, g_ds = wms_deform_attn_backward(
        value, shapes3d[:, :2], lsi, loc[..., :2].contiguous(), attn, ds, grad_output)
    g_dist, g_loc3 = depth_score_sample_backward(dist, shapes3d, lsi, loc, g_ds)
    g_loc3 = g_loc3.clone()
    g_loc3[..., :2] = g_loc3[..., :2] + g_loc2  # F3D:349
    return g_value, g_dist, g_loc3, g_attn


def dfa3d_forward_grid_sample(value, dist, shapes3d, lsi, loc, attn):
    """Independent formulation: 3-D ``grid_sample`` over value (x) dist."""
    B, S, M, Cm = value.shape
    D = dist.shape[-1]
    _, Q, _, L, P, _ = loc.shape
    out = value.new_zeros(B, Q, M, Cm)
    for l in range(L):
        H, W, Dl = (int(x) for x in shapes3d[l])
        start = int(lsi[l])
        v = value[:, start:start + H * W].reshape(B, H, W, M, Cm)
        p = dist[:, start:start + H * W].reshape(B, H, W, M, D)[..., :Dl]
        vol = torch.einsum('bhwmc,bhwmd->bmcdhw', v, p).reshape(B * M, Cm, Dl, H, W)
        grid = (2 * loc[:, :, :, l] - 1).permute(0, 2, 1, 3, 4).reshape(B * M, Q, P, 1, 3)
        smp = F.grid_sample(vol, grid, mode='bilinear', padding_mode='zeros', align_corners=False)
        smp = smp.reshape(B, M, Cm, Q, P)
        a = attn[:, :, :, l].permute(0, 2, 1, 3)  # [B,M,Q,P]
        out = out + torch.einsum('bmcqp,bmqp->bqmc', smp, a)
    return out.reshape(B, Q, M * Cm)
