"""Host-side orchestration of one view-transform level: autograd Functions around the C-ABI kernels.

PyTorch is plumbing here (device memory, streams, the dense library GEMMs of voxel count / pixel count);
every gather / scatter / projection / selection runs in the hand-written kernels of ``csrc/``.
No function in this module has a CPU or eager fallback.
"""
from __future__ import annotations

import ctypes
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
MAX_VIEWS = 128   # csrc/sgc_crossview.cu kMaxViews


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

BF16 = torch.bfloat16
F32 = torch.float32
import os as _os
# backward of the lift as a gather over pixel tiles (csrc/sgc_lift_tiles.cu) instead of the scatter kernel (REDs into a
# zero-filled grad_vg).  Parity-green, but its first version is latency-bound per tile (830 us vs 195 + 58 us of fill at the
# finest ScanNet level): off by default
LIFT_TILES = _os.environ.get('SGC_LIFT_TILES', '0') != '0'
TOPK_MC_MIN = 32768  # levels with more voxels use the many-CTA top-k
UP_BWD_SEPARABLE = True   # False = the one-launch gather form of the upsample backward (kept for its tests)
TOPK_GRID = _os.environ.get('SGC_TOPK_GRID', '1') != '0'   # one-launch grid top-k (round 2); 0 = the round-1 kernels
_TOPK_SCRATCH = {}


def split_cols(x: torch.Tensor, pattern: int) -> torch.Tensor:
    """[R,K] fp32 -> [R,3K] bf16 = (hi|lo|hi) for pattern 0, (hi|hi|lo) for pattern 1 (K-concatenation)."""
    x = x.contiguous()
    R, K = x.shape
    out = torch.empty(R, 3 * K, device=x.device, dtype=BF16)
    call('sgc_split_bf16x3', ptr(x), R, K, K, 1, pattern, ptr(out), stream())
    return out


def split_rows(x: torch.Tensor, group: int, pattern: int) -> torch.Tensor:
    """[G*group, K] fp32 -> [G, 3*group, K] bf16: the three slots are stacked along the ROW (reduction) axis of
    every group of ``group`` rows."""
    x = x.contiguous()
    R, K = x.shape
    out = torch.empty(R // group, 3 * group, K, device=x.device, dtype=BF16)
    call('sgc_split_bf16x3', ptr(x), R, K, K, group, pattern, ptr(out), stream())
    return out


def rows_linear(x: torch.Tensor, wpack: torch.Tensor, N: int, bias: Optional[torch.Tensor] = None, n_cta: int = 0):
    """y [R,N] = x [R,K] @ W^T (+ bias) with ``wpack`` = the packed [N,K] weight (``sgc_rows_gemm_tc``)."""
    R, K = x.shape
    y = torch.empty(R, N, device=x.device, dtype=F32)
    call('sgc_rows_gemm_tc', ptr(x), K, 0, R, K, 1, ptr(wpack), N, 0, 0, ptr(bias), 0, N, ptr(y), N, 0, n_cta, stream())
    return y


def rows_heads_in(x: torch.Tensor, wpack_heads: torch.Tensor, N: int, heads: int = NUM_HEADS, n_cta: int = 0):
    """y [H,R,N]: y[h] = x[:, h*dh:(h+1)*dh] @ W_h^T with ``wpack_heads`` = H packed [N,dh] weights back to back."""
    R, C = x.shape
    dh = C // heads
    y = torch.empty(heads, R, N, device=x.device, dtype=F32)
    call('sgc_rows_gemm_tc', ptr(x), C, dh, R, dh, heads, ptr(wpack_heads), N, 2 * N * dh, 0, None, 0, N, ptr(y), N, R * N,
         n_cta, stream())
    return y


def rows_heads_out(x: torch.Tensor, wpack: torch.Tensor, dh: int, bias: Optional[torch.Tensor] = None, n_cta: int = 0,
                   out: Optional[torch.Tensor] = None):
    """y [R,H*dh]: y[:, h*dh:(h+1)*dh] = x[h] @ W[h*dh:(h+1)*dh]^T (+ bias) with ``wpack`` = the packed [H*dh,K] weight."""
    H, R, K = x.shape
    y = torch.empty(R, H * dh, device=x.device, dtype=F32) if out is None else out
    call('sgc_rows_gemm_tc', ptr(x), K, R * K, R, K, H, ptr(wpack), H * dh, 0, dh, ptr(bias), dh, dh, ptr(y), H * dh, dh,
         n_cta, stream())
    return y


def rows_heads_in_exp(x: torch.Tensor, wpack_exp: torch.Tensor, heads: int = NUM_HEADS, n_cta: int = 0):
    """y [H,R,C]: y[h] = x[:, head h] @ W_h for heads NARROWER than a k-slab (16 wide): every head reads the whole row and its
    weights are zero-extended over all C columns (``wpack_exp`` = the packed [H*C, C] matrix, C rows per head)."""
    R, C = x.shape
    y = torch.empty(heads, R, C, device=x.device, dtype=F32)
    call('sgc_rows_gemm_tc_ex', ptr(x), C, 0, R, C, heads, ptr(wpack_exp), heads * C, 0, C, None, 0, C, ptr(y), C, R * C, n_cta,
         1, 0, stream())
    return y


def rows_heads_out_exp(x: torch.Tensor, wpack_cat: torch.Tensor, bias: Optional[torch.Tensor] = None, n_cta: int = 0,
                       out: Optional[torch.Tensor] = None):
    """y [R,C] = sum_h x[h] @ Wm_h^T with Wm_h the [C,C] weight masked to the output rows of head h (``wpack_cat`` = the packed
    [C, H*C] concatenation along K): the per-head output products of narrow heads as ONE K-concatenated GEMM."""
    H, R, C = x.shape
    y = torch.empty(R, C, device=x.device, dtype=F32) if out is None else out
    call('sgc_rows_gemm_tc_ex', ptr(x), C, R * C, R, H * C, 1, ptr(wpack_cat), C, 0, 0, ptr(bias), 0, C, ptr(y), C, 0, n_cta, 2, H,
         stream())
    return y


_HEAD_MASKS = {}


def _head_mask(dev, C: int, heads: int) -> torch.Tensor:
    """[heads, C] 0/1: channel c belongs to head h."""
    key = (dev, C, heads)
    m = _HEAD_MASKS.get(key)
    if m is None:
        m = _HEAD_MASKS[key] = (torch.arange(C, device=dev) // (C // heads) == torch.arange(heads, device=dev).unsqueeze(1)).to(F32)
    return m


def rows_wgrad(a, b, M, N, R, out, out_strides, *, B=1, lda=None, batch_a=0, ldb=None, batch_b=0, scale=1.0,
               bias_out=None, bias_from=0):
    """``sgc_rows_wgrad_tc``: out[b*ob + m*om + n*on] = scale * sum_r a[b][r, m] * b[b][r, n] (+ column sums of a / b into
    ``bias_out``).  ``out`` / ``bias_out`` may be slices of larger gradient tensors (they are written in place)."""
    scratch = torch.empty(_lib.load().sgc_rows_wgrad_tc_scratch_floats(M, N, R, B), device=a.device, dtype=F32)
    ob, om, on = out_strides
    call('sgc_rows_wgrad_tc', ptr(a), M if lda is None else lda, batch_a, M, ptr(b), N if ldb is None else ldb, batch_b, N,
         R, B, ptr(out), ob, om, on, scale, ptr(bias_out), bias_from, ptr(scratch), stream())
    return out


class WgradGroup:
    """Collects weight-gradient products over the same R voxel rows and issues them as ONE ``sgc_rows_wgrad_group_tc`` launch
    (+ one reduce launch).  ``add`` has the conventions of ``rows_wgrad``; ``linear`` those of ``linear_grads_tc``."""

    def __init__(self, R: int, dev):
        self.R, self.dev, self.jobs, self.keep = R, dev, [], []

    def add(self, a, b, M, N, out, out_strides, *, B=1, lda=None, batch_a=0, ldb=None, batch_b=0, scale=1.0, bias_out=None,
            bias_from=0):
        ob, om, on = out_strides
        self.jobs.append(_lib.WgradJob(ptr(a), M if lda is None else lda, batch_a, M, ptr(b), N if ldb is None else ldb,
                                       batch_b, N, B, ptr(out), ob, om, on, scale, ptr(bias_out), bias_from))
        self.keep.extend((a, b, out, bias_out))
        return out

    def linear(self, g, x, gw=None, gb=None):
        R, N = g.shape
        K = x.shape[1]
        gw = torch.empty(N, K, device=self.dev, dtype=F32) if gw is None else gw
        gb = torch.empty(N, device=self.dev, dtype=F32) if gb is None else gb
        self.add(g, x, N, K, gw, (0, K, 1), bias_out=gb, bias_from=1)
        return gw, gb

    def launch(self):
        lib = _lib.load()
        for i in range(0, len(self.jobs), _lib.MAX_WGRAD_JOBS):
            chunk = self.jobs[i:i + _lib.MAX_WGRAD_JOBS]
            arr = (_lib.WgradJob * len(chunk))(*chunk)
            n = lib.sgc_rows_wgrad_group_scratch_floats(ctypes.cast(arr, ctypes.c_void_p), len(chunk), self.R)
            if n <= 0:
                raise RuntimeError('sgcdet_b200: invalid weight-gradient job table')
            scratch = torch.empty(n, device=self.dev, dtype=F32)
            call('sgc_rows_wgrad_group_tc', ctypes.cast(arr, ctypes.c_void_p), len(chunk), self.R, ptr(scratch), stream())
        self.jobs, self.keep = [], []


def linear_grads_tc(g: torch.Tensor, x: torch.Tensor, gw: Optional[torch.Tensor] = None, gb: Optional[torch.Tensor] = None):
    """(gW [N,K], gb [N]) of y = x W^T + b given g = dL/dy [R,N] and x [R,K] on the own tensor-core kernel; ``gw`` /
    ``gb`` (optional) are the destinations (e.g. row slices of in_proj_weight's gradient)."""
    R, N = g.shape
    K = x.shape[1]
    if gw is None:
        gw = torch.empty(N, K, device=g.device, dtype=F32)
    if gb is None:
        gb = torch.empty(N, device=g.device, dtype=F32)
    rows_wgrad(g, x, N, K, R, gw, (0, K, 1), bias_out=gb, bias_from=1)
    return gw, gb


def pack_weight_tc(w: torch.Tensor) -> torch.Tensor:
    """[N,C] fp32 -> bf16 hi/lo slabs in the shared-memory image of sgc_project_tc_fwd (2*N*C bf16)."""
    N, C = w.shape
    out = torch.empty(2 * N * C, device=w.device, dtype=BF16)
    call('sgc_pack_weight_tc', ptr(w.contiguous()), N, C, ptr(out), stream())
    return out


class FoldWeights(torch.autograd.Function):
    """(Wcat [C+128, C], gbias [128]) of a level from the four projection weights / three small biases of
    ``MSDeformableAttention3D_DFA3D`` (``sgc_fold_wcat``); the backward hands every parameter its own contiguous gradient
    (``sgc_unfold_wcat_grad``), so autograd's accumulation takes them as they are."""

    @staticmethod
    def forward(ctx, wv, wo, wd, wa, bo, bd, ba):
        C = wv.shape[0]
        MP = wd.shape[0]
        wcat = torch.empty(C + 4 * MP, C, device=wv.device, dtype=F32)
        gbias = torch.empty(4 * MP, device=wv.device, dtype=F32)
        call('sgc_fold_wcat', ptr(wv), ptr(wo), ptr(wd), ptr(wa), ptr(bo), ptr(bd), ptr(ba), C, MP, ptr(wcat), ptr(gbias), stream())
        ctx.dims = (C, MP)
        return wcat, gbias

    @staticmethod
    def backward(ctx, gwcat, ggbias):
        C, MP = ctx.dims
        dev = gwcat.device if gwcat is not None else ggbias.device
        if gwcat is None:
            gwcat = torch.zeros(C + 4 * MP, C, device=dev, dtype=F32)
        if ggbias is None:
            ggbias = torch.zeros(4 * MP, device=dev, dtype=F32)
        new = lambda *s: torch.empty(*s, device=dev, dtype=F32)
        gwv, gwo, gwd, gwa, gbo, gbd, gba = new(C, C), new(2 * MP, C), new(MP, C), new(MP, C), new(2 * MP), new(MP), new(MP)
        call('sgc_unfold_wcat_grad', ptr(gwcat.contiguous()), ptr(ggbias.contiguous()), C, MP, ptr(gwv), ptr(gwo), ptr(gwd),
             ptr(gwa), ptr(gbo), ptr(gbd), ptr(gba), stream())
        return gwv, gwo, gwd, gwa, gbo, gbd, gba


class _WeightJobs:
    """Collects the operand preparations of one level and issues them as ONE ``sgc_prepare_weights`` launch."""

    def __init__(self, dev):
        self.dev, self.jobs, self.keep = dev, [], []

    def _add(self, x, out, rpg, pattern, scale, kind):
        assert x.dim() == 2 and x.dtype == F32 and x.is_cuda
        self.jobs.append(_lib.WeightJob(x.data_ptr(), out.data_ptr(), x.stride(0), x.stride(1), x.shape[0], x.shape[1],
                                        rpg, pattern, scale, kind))
        self.keep.append(x)
        return out

    def split_cols(self, x, pattern, scale=1.0):
        return self._add(x, torch.empty(x.shape[0], 3 * x.shape[1], device=self.dev, dtype=BF16), 1, pattern, scale, 0)

    def split_rows(self, x, group, pattern, scale=1.0):
        out = torch.empty(x.shape[0] // group, 3 * group, x.shape[1], device=self.dev, dtype=BF16)
        return self._add(x, out, group, pattern, scale, 0)

    def pack(self, x, scale=1.0):
        return self._add(x, torch.empty(2 * x.numel(), device=self.dev, dtype=BF16), 1, 0, scale, 1)

    def pack_heads_t(self, w, heads, scale=1.0):
        """w [heads*dh, C] -> heads packed [C, dh] matrices (W_h^T, the operand of x_h @ W_h) back to back."""
        assert w.dim() == 2 and w.dtype == F32 and w.is_cuda and w.stride(1) == 1
        C = w.shape[1]
        dh = w.shape[0] // heads
        out = torch.empty(heads * 2 * C * dh, device=self.dev, dtype=BF16)
        for h in range(heads):
            self.jobs.append(_lib.WeightJob(w.data_ptr() + 4 * h * dh * w.stride(0), out.data_ptr() + 2 * h * 2 * C * dh,
                                            1, w.stride(0), C, dh, 1, 0, scale, 1))
        self.keep.append(w)
        return out

    def launch(self):
        for i in range(0, len(self.jobs), _lib.MAX_WEIGHT_JOBS):
            chunk = self.jobs[i:i + _lib.MAX_WEIGHT_JOBS]
            arr = (_lib.WeightJob * len(chunk))(*chunk)
            call('sgc_prepare_weights', ctypes.cast(arr, ctypes.c_void_p), len(chunk), stream())
        self.jobs, self.keep = [], []


class LevelWeights:
    """Every operand image of one level's weights, produced once per step by a single launch (``sgc_prepare_weights``) --
    normally on a side stream, off the critical path of the level: the packed bf16 hi/lo slabs the tcgen05 kernels stream
    (``p_x`` for y = a @ x^T, ``p_x_t`` for the data gradient g @ x) and, for the library fallback of the feature projection
    only, the bf16x3 images of ``wcat``.  Constants for the autograd Functions below (weight gradients are formed from the
    fp32 activations, not from these)."""

    def __init__(self, wcat, w_out, in_w, wo, w1, w2, num_heads: int = NUM_HEADS):
        with torch.no_grad():
            C = w_out.shape[0]
            dh = C // num_heads
            if C % 128 or w1.shape[0] % 128 or dh % 8:
                raise ValueError('sgcdet_b200: the voxel-count layers need embed_dims and feedforward_channels in multiples of '
                                 f'128 and heads at least 8 wide (got {C}, {w1.shape[0]}, {dh})')
            scale = 1.0 / math.sqrt(dh)
            wq, wk, wv = in_w[:C], in_w[C:2 * C], in_w[2 * C:]
            j = _WeightJobs(w_out.device)
            self.wcat = j.split_cols(wcat, 1)          # [N,3C]   x @ Wcat^T   (library fallback of ProjectFeatures)
            self.wcat_t = j.split_cols(wcat.t(), 1)    # [C,3N]   g @ Wcat
            ok = wcat.shape[1] % 32 == 0 and wcat.shape[0] % 32 == 0
            self.wpack = j.pack(wcat) if ok else None
            self.wpack_t = j.pack(wcat.t()) if ok else None
            self.p_w_out, self.p_w_out_t = j.pack(w_out), j.pack(w_out.t())
            self.p_wq, self.p_wq_t = j.pack(wq), j.pack(wq.t())
            self.p_wo, self.p_wo_t = j.pack(wo), j.pack(wo.t())
            self.p_w1, self.p_w1_t = j.pack(w1), j.pack(w1.t())
            self.p_w2, self.p_w2_t = j.pack(w2), j.pack(w2.t())
            # per-head key / value products: heads as wide as a 32-column k-slab (C = 256) are addressed through strided
            # tensor maps; narrower heads (dh = 16 at C = 128) run on the same kernel with zero-extended / K-concatenated
            # weights: 8x redundant MMA work on tiny matrices instead of per-head operand images
            self.heads_tc = dh % 32 == 0
            if self.heads_tc:
                self.p_wk = j.pack(wk, scale)                          # gqv[:, h] = gqt[h] @ (scale Wk_h)^T
                self.p_wv = j.pack(wv)                                 # o[:, h]   = t[h] @ Wv_h^T
                self.p_wk_ht = j.pack_heads_t(wk, num_heads, scale)    # qt[h]     = qv_h @ (scale Wk_h)
                self.p_wv_ht = j.pack_heads_t(wv, num_heads)           # gt[h]     = go_h @ Wv_h
            else:
                hm = _head_mask(w_out.device, C, num_heads)                                          # [H, C]
                self.p_wk_in = j.pack(((wk.t() * scale).unsqueeze(0) * hm.unsqueeze(1)).reshape(num_heads * C, C))
                self.p_wv_in = j.pack((wv.t().unsqueeze(0) * hm.unsqueeze(1)).reshape(num_heads * C, C))
                self.p_wv_out = j.pack((wv.unsqueeze(1) * hm.t().unsqueeze(2)).reshape(C, num_heads * C))
                self.p_wk_out = j.pack(((wk * scale).unsqueeze(1) * hm.t().unsqueeze(2)).reshape(C, num_heads * C))
            j.launch()

    def record_stream(self, s):
        for t in self.__dict__.values():
            if isinstance(t, torch.Tensor):
                t.record_stream(s)


_COUNTERS = {}
_GRAD_STREAMS = {}


GRAD_REDUCER = None   # set by peer.GradAverager: callable(params, grads) -> grads, applied to a parameter group's gradients
GRAD_READY = {}       # data_ptr of a parameter gradient -> event recorded by its PRODUCER right behind the producing launch


def mark_grads_ready(tensors, stream_) -> None:
    """Called by a backward that produces parameter gradients on ``stream_`` while a gradient reducer is installed: the
    reducer waits for this event instead of one recorded when the gradients reach ``OnStream.backward`` -- by then (it is
    among the last nodes autograd executes) unrelated late work has been queued on the same stream (the occupancy head's
    weight gradient, the depth map's layout backward), which held the collective back until the end of the step."""
    ev = torch.cuda.Event()
    ev.record(stream_)
    for t in tensors:
        if t is not None:
            GRAD_READY[t.data_ptr()] = ev


class OnStream(torch.autograd.Function):
    """Identity on parameters, applied under ``with torch.cuda.stream(wstream)``: the aliases' backward node (and the
    AccumulateGrad behind it) belong to ``wstream``, so a backward that PRODUCES a weight gradient on ``wstream`` can
    hand it to autograd without making the calling stream wait for it (autograd orders consumers after the
    producing stream, and joins every leaf stream when the backward pass ends).

    Its backward is also the point where the gradients of a parameter group are final: with scene-batch data parallelism
    (``peer.GradAverager`` installs ``GRAD_REDUCER``) they are averaged over the ranks right here, on a communication stream,
    while the rest of the backward continues -- the averaged tensors are what autograd accumulates."""

    @staticmethod
    def forward(ctx, *ts):
        ctx.params = ts if GRAD_REDUCER is not None else None
        return tuple(t.view_as(t) for t in ts)

    @staticmethod
    def backward(ctx, *gs):
        if GRAD_REDUCER is not None and ctx.params is not None:
            return GRAD_REDUCER(ctx.params, gs)
        return gs


class _Side:
    """Runs the weight/bias-gradient work of a backward on a side stream so that it overlaps the latency-bound
    activation-gradient chain.  With ``wstream`` (the stream the parameters were aliased on, see ``OnStream``) the
    results are never joined into the calling stream; without it a private side stream is used and ``join()`` makes
    the caller wait before the grads are handed to autograd."""

    def __init__(self, dev, wstream=None):
        self.main = torch.cuda.current_stream(dev)
        self.enabled = _os.environ.get('SGC_SIDE_GRADS', '1') != '0'
        self.detached = wstream is not None and self.enabled and wstream != self.main
        if self.detached:
            self.side = wstream
        else:
            key = (dev, self.main.cuda_stream)
            s = _GRAD_STREAMS.get(key)
            if s is None:
                s = _GRAD_STREAMS[key] = torch.cuda.Stream(device=dev)
            self.side = s
        self.out = []

    def run(self, fn, *inputs):
        if not self.enabled:
            return fn()
        self.side.wait_stream(self.main)
        for t in inputs:
            t.record_stream(self.side)
        with torch.cuda.stream(self.side):
            r = fn()
        self.out.extend(r if isinstance(r, tuple) else (r,))
        return r

    def join(self):
        if self.enabled and self.out and not self.detached:
            self.main.wait_stream(self.side)
            for t in self.out:
                t.record_stream(self.main)


def colsum(x: torch.Tensor) -> torch.Tensor:
    """x[R,C] -> [C] column sums (bias gradients), deterministic two-stage reduction in one launch."""
    x = x.contiguous()
    R, C = x.shape
    dev = x.device
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)
    ctr = _COUNTERS.get(key)
    if ctr is None:
        ctr = _COUNTERS[key] = torch.zeros(1, device=dev, dtype=torch.int32)
    out = torch.empty(C, device=dev, dtype=F32)
    scratch = torch.empty(_lib.load().sgc_colsum_scratch_floats(R, C), device=dev, dtype=F32)
    call('sgc_colsum', ptr(x), R, C, ptr(out), ptr(scratch), ptr(ctr), stream())
    return out


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
    def _split_feat(feat, h, w):
        feat = feat.reshape(-1, *feat.shape[-3:])
        V, C, H0, W0 = feat.shape
        S = h * w
        src = feat
        if w != W0 or not feat.is_contiguous():
            src = feat[:, :, :h, :w].contiguous()
            stride = S
        else:
            stride = H0 * W0  # row crop only: the first h*w elements of every channel plane
        acat = torch.empty(V, 3 * C, S, device=feat.device, dtype=BF16)  # (hi|lo|hi) along channels
        call('sgc_split_bf16x3', ptr(src), V * C, S, stride, C, 0, ptr(acat), stream())
        return acat

    @staticmethod
    def forward(ctx, feat: torch.Tensor, h: int, w: int, wcat: torch.Tensor, lw=None):
        # feat is [V,C,H0,W0] or the reference's [1,V,C,H0,W0]; taking the 5-D leaf directly keeps autograd from
        # materialising a zero-filled copy for the select() view on the way back
        ctx.feat_shape = tuple(feat.shape)
        feat = feat.reshape(-1, *feat.shape[-3:])
        V, C, H0, W0 = feat.shape
        S = h * w
        N = wcat.shape[0]
        ctx.dims = (V, C, H0, W0, h, w)
        ctx.lw = lw
        use_tc = (_os.environ.get('SGC_TC_PROJECT', '1') != '0' and w == W0 and feat.is_contiguous()
                  and (H0 * W0 * 4) % 16 == 0 and C % 32 == 0 and N % 32 == 0 and N <= 512)
        if use_tc:
            # own tcgen05 kernel: TMA-loads the fp32 NCHW map, splits to bf16 hi/lo in shared memory, accumulates in
            # TMEM and TMA-stores channel-last fp32 (csrc/sgc_project_tc.cu); the split never touches HBM
            wpack = lw.wpack if lw is not None and getattr(lw, 'wpack', None) is not None else pack_weight_tc(wcat)
            vg = torch.empty(V, S, N, device=feat.device, dtype=F32)
            call('sgc_project_tc_fwd', ptr(feat), C * H0 * W0, H0 * W0, V, C, S, ptr(wpack), N, ptr(vg), stream())
            ctx.save_for_backward(feat, wcat)
            ctx.have_acat = False
            return vg
        acat = ProjectFeatures._split_feat(feat, h, w)
        bcat = lw.wcat if lw is not None else split_cols(wcat, 1)  # [N, 3C]
        vg = torch.bmm(acat.transpose(1, 2), bcat.t().unsqueeze(0).expand(V, -1, -1), out_dtype=F32)  # [V,S,N]
        ctx.save_for_backward(acat, wcat)
        ctx.have_acat = True
        return vg

    @staticmethod
    def backward(ctx, gvg: torch.Tensor):
        acat, wcat = ctx.saved_tensors
        V, C, H0, W0, h, w = ctx.dims
        S = h * w
        N = wcat.shape[0]
        if (not ctx.have_acat) and _os.environ.get('SGC_TC_BWD', '1') != '0' and C <= 256 and N % 128 == 0:
            # own tcgen05 kernels for both gradients: gvg and feat are read as fp32 and split in shared memory
            feat = acat
            gvg = gvg.contiguous()
            gfeat = gw = None
            lw = ctx.lw
            if ctx.needs_input_grad[0]:
                wpack_t = lw.wpack_t if lw is not None and getattr(lw, 'wpack_t', None) is not None \
                    else pack_weight_tc(wcat.t().contiguous())
                gfeat = torch.empty(V, C, H0, W0, device=gvg.device, dtype=F32)
                if h != H0:
                    gfeat[:, :, h:].zero_()
                call('sgc_project_tc_bwd_data', ptr(gvg), V, S, N, ptr(wpack_t), C, ptr(gfeat), H0 * W0, stream())
                gfeat = gfeat.view(ctx.feat_shape)
            if ctx.needs_input_grad[3]:
                gw = torch.empty(N, C, device=gvg.device, dtype=F32)
                scratch = torch.empty(_lib.load().sgc_project_tc_wgrad_scratch_floats(N, C), device=gvg.device, dtype=F32)
                call('sgc_project_tc_wgrad', ptr(gvg), ptr(feat), H0 * W0, V, S, N, C, ptr(gw), ptr(scratch), stream())
            return gfeat, None, None, gw, None
        if not ctx.have_acat:
            acat = ProjectFeatures._split_feat(acat, h, w) if ctx.needs_input_grad[3] else None
        gvg = gvg.contiguous()
        gfeat = gw = gcat = None
        if ctx.needs_input_grad[0]:
            gcat = split_cols(gvg.view(V * S, N), 0).view(V, S, 3 * N)
            wk = ctx.lw.wcat_t if ctx.lw is not None else split_cols(wcat.t(), 1)  # [C, 3N]
            if w == W0:
                # write straight into the padded NCHW gradient; only the cropped rows need zeroing
                gfeat = torch.empty(V, C, H0, W0, device=gvg.device, dtype=F32)
                if h != H0:
                    gfeat[:, :, h:].zero_()
                torch.bmm(wk.unsqueeze(0).expand(V, -1, -1), gcat.transpose(1, 2), out_dtype=F32,
                          out=gfeat.view(V, C, H0 * W0)[:, :, :S])
            else:
                g = torch.bmm(wk.unsqueeze(0).expand(V, -1, -1), gcat.transpose(1, 2), out_dtype=F32)
                gfeat = g.new_zeros(V, C, H0, W0)
                gfeat[:, :, :h, :w] = g.view(V, C, h, w)
        if ctx.needs_input_grad[3]:
            # gw[n,c] = sum_{v,s} g[v,s,n] f[v,c,s] = g_hi f_hi + g_hi f_lo + g_lo f_hi, straight from the
            # already split operands (strided views, no copies): acat = (f_hi | f_lo | f_hi) along channels
            if gcat is None:
                gcat = split_cols(gvg.view(V * S, N), 0).view(V, S, 3 * N)
            x = torch.bmm(gcat[:, :, :N].transpose(1, 2), acat[:, :2 * C].transpose(1, 2), out_dtype=F32)  # [V,N,2C]
            y = torch.bmm(gcat[:, :, N:2 * N].transpose(1, 2), acat[:, :C].transpose(1, 2), out_dtype=F32)  # [V,N,C]
            xs, ys = x.sum(0), y.sum(0)
            gw = xs[:, :C] + xs[:, C:] + ys
        if gfeat is not None:
            gfeat = gfeat.view(ctx.feat_shape)
        return gfeat, None, None, gw, None


# ----------------------------------------------------------------------------------------------
# lift
# ----------------------------------------------------------------------------------------------

class Lift(torch.autograd.Function):
    """sgc_lift_fwd / sgc_lift_bwd (see csrc/sgc_lift.cu)."""

    @staticmethod
    def forward(ctx, vg, dist, vbias, gbias, pl: PairList, H: int, W: int, bwd_stream=None, join_stream=None):
        """``bwd_stream``: the stream the producers of vg / dist / vbias / gbias ran on (``DenseHead.prepare`` on a
        side stream).  Autograd replays those producers' backward nodes on that stream, so the (large) backward kernel
        is issued there as well and the main stream continues with the next level's per-voxel chain."""
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
        ctx.bwd_stream = bwd_stream if _os.environ.get('SGC_SIDE_LIFT_BWD', '1') != '0' else None
        # when bwd_stream is NOT the stream the producers of vg / dist ran on, that producer stream (join_stream) has to
        # wait for the backward kernel before autograd replays the producers' backward nodes on it
        ctx.join_stream = join_stream if ctx.bwd_stream is not None else None
        ctx.mark_non_differentiable(samp)
        ctx.set_materialize_grads(False)
        return slots, samp

    @staticmethod
    def backward(ctx, gslots, _gsamp):
        vg, dist, vbias, samp = ctx.saved_tensors
        pl = ctx.pl
        S, H, W, D, C = ctx.dims
        ld = vg.shape[-1]
        gslots = gslots.contiguous()
        cur = torch.cuda.current_stream(vg.device)
        side = ctx.bwd_stream if ctx.bwd_stream is not None and ctx.bwd_stream != cur else None
        tiles = LIFT_TILES and S == H * W
        lib = _lib.load()

        def _zeros():
            return (torch.zeros_like(vg), torch.zeros_like(dist), torch.zeros_like(vbias),
                    torch.zeros(G_CH, device=vg.device, dtype=torch.float32))
        prezero = (not tiles) and _os.environ.get('SGC_PREZERO', '1') != '0'
        with torch.cuda.stream(side if side is not None else cur):
            # scatter kernel only: the accumulation targets are zero-filled BEFORE the side stream joins the voxel chain (the
            # fills, 290 MB at the finest level, depend on nothing); the tile kernel writes every row once and needs none
            if prezero:
                gvg, gdist, gvb, ggb = _zeros()
        if side is not None:
            # every consumer of the four gradients is a backward node of the side stream, so main never waits
            side.wait_stream(cur)
            for t in (gslots, samp, pl.pair_vq, pl.n_pairs, pl.ref_cam):
                t.record_stream(side)
        with torch.cuda.stream(side if side is not None else cur):
            base = ptr(vg)
            if tiles:
                # gather over pixel tiles (csrc/sgc_lift_tiles.cu): no zero fill, no reductions into grad_vg
                V = vg.shape[0]
                gvg, gdist = torch.empty_like(vg), torch.empty_like(dist)
                gvb = torch.empty_like(vbias)
                ggb = torch.empty(G_CH, device=vg.device, dtype=torch.float32)
                ws = torch.empty(lib.sgc_lift_bwd_tiles_workspace_bytes(pl.cap, V, H, W, C), device=vg.device, dtype=torch.uint8)
                gbase = ptr(gvg)
                call('sgc_lift_bwd_tiles', base, ld, base + 4 * C, ld, ptr(dist), ptr(vbias), ptr(pl.pair_vq), ptr(pl.n_pairs),
                     pl.cap, ptr(pl.ref_cam), ptr(samp), ptr(gslots), V, S, H, W, D, pl.Q, C, gbase, gbase + 4 * C, ptr(gdist),
                     ptr(gvb), ptr(ggb), ptr(ws), stream())
            else:
                if not prezero:
                    gvg, gdist, gvb, ggb = _zeros()
                gbase = ptr(gvg)
                scratch = torch.empty(lib.sgc_lift_bwd_scratch_floats(pl.cap, C), device=vg.device, dtype=torch.float32)
                call('sgc_lift_bwd', base, ld, base + 4 * C, ld, ptr(dist), ptr(vbias), ptr(pl.pair_vq), ptr(pl.n_pairs),
                     pl.cap, ptr(pl.ref_cam), ptr(samp), ptr(gslots), S, H, W, D, pl.Q, C,
                     gbase, gbase + 4 * C, ptr(gdist), ptr(gvb), ptr(ggb), ptr(scratch), stream())
        if side is not None and ctx.join_stream is not None and ctx.join_stream != side:
            ctx.join_stream.wait_stream(side)
            for t in (gvg, gdist, gvb, ggb):
                t.record_stream(ctx.join_stream)
        return gvg, gdist, gvb, ggb, None, None, None, None, None


# ----------------------------------------------------------------------------------------------
# fused encoder layer over voxel rows
# ----------------------------------------------------------------------------------------------

def rowop_fwd(x, R, N, *, bias=None, relu=False, mask=None, mscale=1.0, rowscale=None, residual=None, ln=None,
              in_heads=0, split_heads=0, want_y=True, want_split=True, rowcount=None):
    """``sgc_rowop_fwd``: fused epilogue of a GEMM output ``x`` ([R,N], or head-major [H,R,N/H] with in_heads=H).
    Returns (y [R,N] fp32 | None, ysplit bf16x3 | None, (pre, mean, rstd) | None)."""
    dev = x.device
    y = torch.empty(R, N, device=dev, dtype=F32) if want_y else None
    ys = torch.empty(R, 3 * N, device=dev, dtype=BF16) if want_split else None
    saved = None
    a = _lib.RowopFwdArgs()
    a.x, a.bias, a.mask, a.rowscale, a.residual = ptr(x), ptr(bias), ptr(mask), ptr(rowscale), ptr(residual)
    a.rowcount = ptr(rowcount)
    if ln is not None:
        gamma, beta, eps = ln
        saved = (torch.empty(R, N, device=dev, dtype=F32), torch.empty(R, device=dev, dtype=F32),
                 torch.empty(R, device=dev, dtype=F32))
        a.gamma, a.beta, a.eps = ptr(gamma), ptr(beta), eps
        a.pre, a.mean, a.rstd = ptr(saved[0]), ptr(saved[1]), ptr(saved[2])
    a.y, a.ysplit = ptr(y), ptr(ys)
    a.mscale, a.R, a.N, a.relu, a.in_heads, a.split_heads = mscale, R, N, int(relu), in_heads, split_heads
    call('sgc_rowop_fwd', ctypes.byref(a), stream())
    return y, ys, saved


def rowop_bwd(g, R, N, *, g2=None, ln=None, mask=None, mscale=1.0, gate=None, gscale=1.0, rowscale=None,
              in_heads=0, split_heads=0, want_gpre=False, want_gx=True, want_split=True, rowcount=None):
    """``sgc_rowop_bwd``.  ``ln`` = (pre, mean, rstd, gamma).  Returns (gx, gxsplit, gpre, partial)."""
    dev = g.device
    gx = torch.empty(R, N, device=dev, dtype=F32) if want_gx else None
    gs = torch.empty(R, 3 * N, device=dev, dtype=BF16) if want_split else None
    gpre = torch.empty(R, N, device=dev, dtype=F32) if want_gpre else None
    partial = None
    a = _lib.RowopBwdArgs()
    a.g, a.g2, a.mask, a.gate, a.rowscale = ptr(g), ptr(g2), ptr(mask), ptr(gate), ptr(rowscale)
    a.rowcount = ptr(rowcount)
    if ln is not None:
        pre, mean, rstd, gamma = ln
        partial = torch.empty(_lib.load().sgc_layernorm_bwd_scratch_floats(R, N), device=dev, dtype=F32)
        a.pre, a.mean, a.rstd, a.gamma, a.partial = ptr(pre), ptr(mean), ptr(rstd), ptr(gamma), ptr(partial)
    a.gpre, a.gx, a.gxsplit = ptr(gpre), ptr(gx), ptr(gs)
    a.mscale, a.gscale, a.R, a.N, a.in_heads, a.split_heads = mscale, gscale, R, N, in_heads, split_heads
    call('sgc_rowop_bwd', ctypes.byref(a), stream())
    return gx, gs, gpre, partial


class DropoutMasks:
    """Keep-masks of all dropouts of a step in one launch (``sgc_dropout_masks``: Philox4x32-10, the step counter lives on the
    device and advances with every launch, so CUDA-graph replays draw fresh masks).  One instance per call site."""

    def __init__(self, device, seed: Optional[int] = None):
        self.device = torch.device(device)
        self.state = torch.zeros(2, device=self.device, dtype=torch.int64)
        # torch's seed at creation plus a per-instance salt: two call sites never share a stream of masks
        DropoutMasks._instances += 1
        s = (torch.initial_seed() if seed is None else int(seed)) + 0x9E3779B97F4A7C15 * DropoutMasks._instances
        self.seed = ((s + 2 ** 63) % 2 ** 64) - 2 ** 63   # as a signed 64-bit integer for the C ABI

    _instances = 0

    def draw(self, specs):
        """specs: [(rows, width, p)] -> list of uint8 [rows, width] keep-masks (None where p == 0)."""
        todo = [(i, r, w, p) for i, (r, w, p) in enumerate(specs) if p > 0 and r * w > 0]
        out = [None] * len(specs)
        if not todo:
            return out
        sizes = [(r * w + 15) // 16 * 16 for _, r, w, _ in todo]
        buf = torch.empty(sum(sizes), device=self.device, dtype=torch.uint8)
        jobs = (_lib.MaskJob * len(todo))()
        off = 0
        for k, ((i, r, w, p), sz) in enumerate(zip(todo, sizes)):
            m = buf[off:off + r * w].view(r, w)
            jobs[k].out, jobs[k].n, jobs[k].keep = ptr(m), r * w, 1.0 - p
            out[i] = m
            off += sz
        call('sgc_dropout_masks', ctypes.addressof(jobs), len(todo), self.seed, ptr(self.state), stream())
        return out


def _ln_params(partial, R, N):
    gg, gb = torch.empty(N, device=partial.device, dtype=F32), torch.empty(N, device=partial.device, dtype=F32)
    call('sgc_layernorm_bwd_params', ptr(partial), R, N, ptr(gg), ptr(gb), stream())
    return gg, gb


def rows_headscale(x, s_, heads: int, smin: float, bias=None, want_count: bool = False):
    """``sgc_rows_headscale``: y[r,c] = x[r,c] / max(s[r, head of c], smin) (+ bias[c]); with ``want_count`` (heads == 1)
    also the int32 view counts (int)s[r].  The finishing step after an exchange of view sharding."""
    R, C = x.shape
    y = torch.empty(R, C, device=x.device, dtype=F32)
    cnt = torch.empty(R, device=x.device, dtype=torch.int32) if want_count else None
    call('sgc_rows_headscale', ptr(x), ptr(s_), heads, smin, ptr(bias), R, C, ptr(y), ptr(cnt), stream())
    return (y, cnt) if want_count else y


class EncoderLayerRows(torch.autograd.Function):
    """One VoxFormerLayer over the selected voxel rows (encoder.py:262-340 with operation_order cross_attn, norm, ffn,
    norm): masked mean over views -> output_proj -> 8-head attention pooling over views (DCA:815-837) -> LayerNorm ->
    FFN (+identity) -> LayerNorm, as ONE fixed sequence of launches: the own tcgen05 GEMMs (``sgc_rows_gemm_tc``, bias in the
    epilogue) alternate with the cross-view kernels and fused row kernels (``sgc_rowop_fwd/bwd``) that carry ReLU, dropout
    mask, row mask, residual and LayerNorm.  The backward is written out by hand; all seven weight-gradient products of the
    layer are ONE grouped launch (``sgc_rows_wgrad_group_tc``) on ``wstream`` = (stream of the attention-block parameters,
    stream of the FFN / norm parameters), see ``OnStream``.

    ``masks`` = (mask_attn, mask_ffn1, mask_ffn2) uint8 keep-masks or None (eval / p = 0), ``drops`` the matching p.

    ``coll`` (view sharding, ``parallel.ViewShardExchange``): ``slots`` / ``pl`` hold only the views this rank owns;
    every statistic over views becomes (local partial, written straight into peer memory) -> one all-reduce launch ->
    (local finish).  Exchanged per level: sums + counts [Q,C+1], score maxima [Q,8], the per-head value products of the
    partial softmax sums together with the normalisers [Q,C+8] (the per-head projection is applied to the PARTIAL sums:
    8x fewer bytes on the links than the [8,Q,C] sums themselves), and in the backward the normaliser dot [Q,8] and the
    query gradient after its per-head projection [Q,C]."""

    @staticmethod
    def forward(ctx, slots, pl: PairList, w_out, b_out, in_w, in_b, wo, bo, w1, b1, w2, b2, g1, be1, g2, be2,
                lw, wstream, eps1, eps2, masks, drops, coll=None):
        Q, V = pl.Q, pl.V
        C = slots.shape[1]
        H = NUM_HEADS
        dh = C // H
        Fh = w1.shape[0]
        dev = slots.device
        wide = lw.heads_tc
        bq, bv = in_b[:C], in_b[2 * C:]
        m0, m1, m2 = masks if masks is not None else (None, None, None)
        s0, s1, s2 = (1.0 / (1.0 - p) if m is not None else 1.0 for m, p in zip((m0, m1, m2), drops))

        def heads_in(x, p_wide, p_narrow):      # y[h] = x[:, head h] @ W_h            -> [H,Q,C]
            return rows_heads_in(x, p_wide, C, H) if wide else rows_heads_in_exp(x, p_narrow, H)

        def heads_out(x, p_wide, p_narrow, bias=None, out=None):   # y[:, head h] = x[h] @ W_h^T (+ bias)  -> [Q,C]
            return rows_heads_out(x, p_wide, dh, bias, out=out) if wide else rows_heads_out_exp(x, p_narrow, bias, out=out)

        cnt = pl.count
        if coll is None:
            mean = torch.empty(Q, C, device=dev, dtype=F32)
            call('sgc_crossview_mean_fwd_split', ptr(slots), ptr(pl.pair_index), V, Q, C, ptr(mean), None, stream())
        else:
            # exchange 1: sums over the local views + local view counts -> mean over ALL views, global counts
            call('sgc_crossview_sum_fwd', ptr(slots), ptr(pl.pair_index), V, Q, C, ptr(coll.view((Q, C))), stream())
            coll.view((Q,), Q * C).copy_(pl.count)
            red = coll.reduce(Q * C + Q, 'sum')
            mean, cnt = rows_headscale(red[:Q * C].view(Q, C), red[Q * C:], 1, 1.0, want_count=True)
        g = rows_linear(mean, lw.p_w_out, C, b_out)
        qv = rows_linear(g, lw.p_wq, C, bq)
        qt = heads_in(qv, getattr(lw, 'p_wk_ht', None), getattr(lw, 'p_wk_in', None))          # scale Wk_h folded in
        t = torch.empty(H, Q, C, device=dev, dtype=F32)
        alpha = torch.empty(pl.cap, H, device=dev, dtype=F32)
        ssm = None
        if coll is None:
            call('sgc_crossview_attn_fwd_split', ptr(qt), ptr(slots), ptr(pl.pair_index), V, Q, C, ptr(t), ptr(alpha), None,
                 stream())
            o2 = heads_out(t, getattr(lw, 'p_wv', None), getattr(lw, 'p_wv_out', None), bv)
        else:
            # exchange 2: score maxima; exchange 3: the partial softmax sums -- the log-sum-exp merge without a rescale pass.
            # t = this rank's UNNORMALISED partial sum_v e_v s_v, alpha = e (both finished in the backward with ssm)
            sc = torch.empty(pl.cap, H, device=dev, dtype=F32)
            call('sgc_cvs_scores', ptr(qt), ptr(slots), ptr(pl.pair_index), V, Q, C, ptr(sc), ptr(coll.view((Q, H))), stream())
            m = coll.reduce(Q * H, 'max')
            call('sgc_cvs_accum', ptr(sc), ptr(m), ptr(slots), ptr(pl.pair_index), V, Q, C, ptr(alpha),
                 ptr(coll.view((Q, H), Q * C)), ptr(t), stream())
            heads_out(t, getattr(lw, 'p_wv', None), getattr(lw, 'p_wv_out', None), out=coll.view((Q, C)))
            red = coll.reduce(Q * C + Q * H, 'sum')
            ssm = red[Q * C:].view(Q, H)
            o2 = rows_headscale(red[:Q * C].view(Q, C), ssm, H, 1e-30, bv)
        # rows no view sees are zeroed (DCA:819-835) straight from the per-voxel view count
        x1, _, ln1 = rowop_fwd(rows_linear(o2, lw.p_wo, C), Q, C, bias=bo, mask=m0, mscale=s0, rowcount=cnt,
                               ln=(g1, be1, eps1), want_split=False)
        hdn, _, _ = rowop_fwd(rows_linear(x1, lw.p_w1, Fh), Q, Fh, bias=b1, relu=True, mask=m1, mscale=s1, want_split=False)
        y, _, ln2 = rowop_fwd(rows_linear(hdn, lw.p_w2, C), Q, C, bias=b2, mask=m2, mscale=s2, residual=x1,
                              ln=(g2, be2, eps2), want_split=False)
        ctx.save_for_backward(slots, mean, g, qv, qt, t, alpha, o2, x1, hdn, *ln1, *ln2, g1, g2)
        ctx.pl, ctx.lw, ctx.wstream = pl, lw, wstream
        ctx.coll, ctx.ssm, ctx.cnt = coll, ssm, cnt
        ctx.masks, ctx.scales = (m0, m1, m2), (s0, s1, s2)
        return y

    @staticmethod
    def backward(ctx, gy):
        (slots, mean, g, qv, qt, t, alpha, o2, x1, hdn, pre1, mean1, rstd1, pre2, mean2, rstd2, g1, g2) = ctx.saved_tensors
        pl, lw = ctx.pl, ctx.lw
        coll, ssm, cnt = ctx.coll, ctx.ssm, ctx.cnt
        m0, m1, m2 = ctx.masks
        s0, s1, s2 = ctx.scales
        Q, V = pl.Q, pl.V
        C = slots.shape[1]
        H = NUM_HEADS
        dh = C // H
        Fh = hdn.shape[1]
        dev = slots.device
        scale = 1.0 / math.sqrt(dh)
        wide = lw.heads_tc
        ws_attn, ws_ffn = ctx.wstream if ctx.wstream is not None else (None, None)
        side = _Side(dev, ws_attn)     # attention-block parameters
        side_f = _Side(dev, ws_ffn)    # FFN + norms
        gy = gy.contiguous()

        def heads_in(x, p_wide, p_narrow):
            return rows_heads_in(x, p_wide, C, H) if wide else rows_heads_in_exp(x, p_narrow, H)

        def heads_out(x, p_wide, p_narrow, out=None):
            return rows_heads_out(x, p_wide, dh, out=out) if wide else rows_heads_out_exp(x, p_narrow, out=out)

        # norm 2 + dropout of the second FFN layer; gpre2 also flows into the identity branch
        gf, _, gpre2, part2 = rowop_bwd(gy, Q, C, ln=(pre2, mean2, rstd2, g2), mask=m2, mscale=s2, want_gpre=True,
                                        want_split=False)
        g_g2, g_be2 = side_f.run(lambda: _ln_params(part2, Q, C), part2)
        ghdn = rows_linear(gf, lw.p_w2_t, Fh)                                                           # [Q,F]
        # hdn = relu(.)*mask1*s1, so the ReLU gate and the dropout mask together are (hdn > 0)
        gh, _, _, _ = rowop_bwd(ghdn, Q, Fh, gate=hdn, gscale=s1, want_split=False)
        gx1_raw = rows_linear(gh, lw.p_w1_t, C)                                                         # [Q,C]
        gout, _, _, part1 = rowop_bwd(gx1_raw, Q, C, g2=gpre2, ln=(pre1, mean1, rstd1, g1), mask=m0, mscale=s0,
                                      rowcount=cnt, want_split=False)
        g_g1, g_be1 = side_f.run(lambda: _ln_params(part1, Q, C), part1)
        go2 = rows_linear(gout, lw.p_wo_t, C)                                                           # [Q,C]
        gt = heads_in(go2, getattr(lw, 'p_wv_ht', None), getattr(lw, 'p_wv_in', None))                  # [H,Q,C]
        gscore = torch.empty(pl.cap, H, device=dev, dtype=F32)
        gqt = torch.empty(H, Q, C, device=dev, dtype=F32)
        go2_n = None
        if coll is None:
            call('sgc_crossview_attn_bwd_qt_split', ptr(slots), ptr(alpha), ptr(pl.pair_index), V, Q, C, ptr(gt), ptr(gscore),
                 ptr(gqt), None, stream())
            gqv = heads_out(gqt, getattr(lw, 'p_wk', None), getattr(lw, 'p_wk_out', None))
        else:
            # exchange 4: the softmax-normaliser dot D[q,h] = sum over ALL views of alpha g_alpha; exchange 5: the query
            # gradient, after its per-head projection (linear, so it applies to the partial sums).  `alpha` arrives as e.
            a_n = torch.empty(pl.cap, H, device=dev, dtype=F32)
            galpha = torch.empty(pl.cap, H, device=dev, dtype=F32)
            call('sgc_cvs_bwd_dot', ptr(slots), ptr(alpha), ptr(ssm), ptr(pl.pair_index), V, Q, C, ptr(gt), ptr(a_n), ptr(galpha),
                 ptr(coll.view((Q, H))), stream())
            dsum = coll.reduce(Q * H, 'sum')
            call('sgc_cvs_bwd_qt', ptr(slots), ptr(a_n), ptr(galpha), ptr(dsum), ptr(pl.pair_index), V, Q, C, ptr(gscore), ptr(gqt),
                 stream())
            alpha = a_n
            heads_out(gqt, getattr(lw, 'p_wk', None), getattr(lw, 'p_wk_out', None), out=coll.view((Q, C)))
            gqv = coll.reduce(Q * C, 'sum').view(Q, C)
            # the value-weight gradient pairs this rank's unnormalised partial t with go2 / ssm (per head): sum over ranks =
            # go_h^T (sum_ranks t / ssm)
            go2_n = rows_headscale(go2, ssm, H, 1e-30)
        gg = rows_linear(gqv, lw.p_wq_t, C)
        gmean = rows_linear(gg, lw.p_w_out_t, C)
        gslots = torch.empty_like(slots)
        if coll is None:
            call('sgc_crossview_attn_bwd_slots', ptr(qt), ptr(alpha), ptr(gscore), ptr(pl.pair_index), V, Q, C, ptr(gt),
                 ptr(gmean), ptr(gslots), stream())
        else:
            call('sgc_cvs_bwd_slots', ptr(qt), ptr(alpha), ptr(gscore), ptr(pl.pair_index), V, Q, C, ptr(gt), ptr(gmean),
                 ptr(cnt), ptr(gslots), stream())

        def _group():
            # the three in-projection gradients are written straight into in_proj_weight's / in_proj_bias's gradients (rows
            # [0,C) query, [C,2C) key, [2C,3C) value; the key bias gradient is identically zero: it cancels in the softmax)
            G = WgradGroup(Q, dev)
            gw_in = torch.empty(3 * C, C, device=dev, dtype=F32)
            gb_in = torch.zeros(3 * C, device=dev, dtype=F32)
            w2g = G.linear(gf, hdn)
            w1g = G.linear(gh, x1)
            wog = G.linear(gout, o2)
            # g_wv[h*dh + d, c] = sum_q t[h][q, c] go2[q, h*dh + d];  g_bv = column sums of go2
            if coll is not None:   # partial over this rank's views (summed over the ranks after the backward)
                G.add(t, go2_n, C, dh, gw_in[2 * C:], (dh * C, 1, C), B=H, lda=C, batch_a=Q * C, ldb=C, batch_b=dh)
                gb_in[2 * C:].copy_(colsum(go2))
            else:
                G.add(t, go2, C, dh, gw_in[2 * C:], (dh * C, 1, C), B=H, lda=C, batch_a=Q * C, ldb=C, batch_b=dh,
                      bias_out=gb_in[2 * C:], bias_from=2)
            # g_wk[h*dh + d, c] = scale * sum_q gqt[h][q, c] qv[q, h*dh + d]
            G.add(gqt, qv, C, dh, gw_in[C:2 * C], (dh * C, 1, C), B=H, lda=C, batch_a=Q * C, ldb=C, batch_b=dh, scale=scale)
            G.linear(gqv, g, gw_in[:C], gb_in[:C])
            woutg = G.linear(gg, mean)
            G.launch()
            return w2g + w1g + wog + woutg + (gw_in, gb_in)
        g_w2, g_b2, g_w1, g_b1, g_wo, g_bo, g_wout, g_bout, g_in_w, g_in_b = side.run(
            _group, gf, hdn, gh, x1, gout, o2, t, go2, gqt, qv, gqv, g, gg, mean, *((go2_n,) if coll is not None else ()))
        if side.detached and side_f.detached and side_f.side != side.side:
            # the FFN parameters are aliased on the second weight stream: it only has to follow the grouped launch
            side_f.side.wait_stream(side.side)
            for t_ in (g_w2, g_b2, g_w1, g_b1):
                t_.record_stream(side_f.side)
        side.join()
        side_f.join()
        if GRAD_REDUCER is not None:
            mark_grads_ready((g_wout, g_bout, g_in_w, g_in_b, g_wo, g_bo), side.side if side.detached else side.main)
            mark_grads_ready((g_w1, g_b1, g_w2, g_b2, g_g1, g_be1, g_g2, g_be2), side_f.side if side_f.detached else side_f.main)
        return (gslots, None, g_wout, g_bout, g_in_w, g_in_b, g_wo, g_bo, g_w1, g_b1, g_w2, g_b2, g_g1, g_be1, g_g2, g_be2,
                None, None, None, None, None, None, None)


# ----------------------------------------------------------------------------------------------
# sparse volume construction
# ----------------------------------------------------------------------------------------------

class UpsampleOcc(torch.autograd.Function):
    """(up [2X,2Y,2Z,C], occ [8XYZ]) = trilinear x2 of a channel-last volume + Linear(C,1)+Sigmoid
    (AdaptiveSparseHead.py:64-71)."""

    @staticmethod
    def forward(ctx, vol, w_occ, b_occ, wstream=None):
        ctx.wstream = wstream
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
        side = _Side(dev, ctx.wstream)
        detached = side.detached and gocc is not None
        scratch = None
        if UP_BWD_SEPARABLE:
            scratch = torch.empty(_lib.load().sgc_upsample2x_occ_bwd_scratch_floats(X, Y, Z, C), device=dev, dtype=torch.float32)
        call('sgc_upsample2x_occ_bwd', ptr(vol), X, Y, Z, C, ptr(w_occ), ptr(occ), ptr(gup),
             ptr(gocc.contiguous()) if gocc is not None else None, ptr(gpre), ptr(gin),
             None if detached else ptr(gw), ptr(gb), ptr(scratch), stream())
        if detached:   # the weight gradient (a full pass over the upsampled volume) leaves the voxel chain
            def _gw():
                g = torch.zeros(C, device=dev, dtype=torch.float32)
                call('sgc_upsample2x_occ_gradw', ptr(vol), X, Y, Z, C, ptr(gpre), ptr(g), stream())
                return g
            gw = side.run(_gw, vol, gpre)
        return gin, gw.view_as(w_occ), gb, None


def topk_select(occ: torch.Tensor, k: int):
    """``topk_wo_grad`` (AdaptiveSparseHead.py:9-13) + ``nonzero`` (DenseHead.py:66).  Returns
    (sel [k] int32 ascending, mask [N] uint8).  Deterministic: ties -> lower index."""
    N = occ.numel()
    sel = torch.empty(k, device=occ.device, dtype=torch.int32)
    mask = torch.empty(N, device=occ.device, dtype=torch.uint8)
    lib = _lib.load()
    if TOPK_GRID and 0 < k and N <= lib.sgc_topk_grid_max_n():
        # one launch of a few co-operating CTAs (csrc/sgc_volume.cu topk_select_grid_kernel); its 2 KB of scratch is
        # zero-filled once per (device, stream) and kept consistent by the kernel itself
        dev = occ.device
        key = (dev, torch.cuda.current_stream(dev).cuda_stream)
        scratch = _TOPK_SCRATCH.get(key)
        if scratch is None:
            scratch = _TOPK_SCRATCH[key] = torch.zeros(lib.sgc_topk_grid_scratch_bytes() // 8, device=dev, dtype=torch.int64)
        call('sgc_topk_select_grid', ptr(occ.detach()), N, k, ptr(sel), ptr(mask), ptr(scratch), stream())
    elif N > TOPK_MC_MIN and k > 0:
        # large levels ("-L" configs): many-CTA radix select instead of one CTA streaming over the scores
        scratch = torch.empty(_lib.load().sgc_topk_scratch_ints(N), device=occ.device, dtype=torch.int32)
        call('sgc_topk_select_mc', ptr(occ.detach()), N, k, ptr(sel), ptr(mask), ptr(scratch), stream())
    else:
        call('sgc_topk_select', ptr(occ.detach()), N, k, ptr(sel), ptr(mask), stream())
    return sel, mask


class OccLoss(torch.autograd.Function):
    """``AdaptiveSparseHead.occ_loss`` (AdaptiveSparseHead.py:100-103): 0.5 * mean(BCELoss(p, t)) as one launch each way."""

    @staticmethod
    def forward(ctx, p, t):
        p, t = p.contiguous(), t.contiguous()
        loss = torch.empty((), device=p.device, dtype=F32)
        call('sgc_occ_loss_fwd', ptr(p), ptr(t), p.numel(), ptr(loss), stream())
        ctx.save_for_backward(p, t)
        return loss

    @staticmethod
    def backward(ctx, g):
        p, t = ctx.saved_tensors
        gp = torch.empty_like(p)
        call('sgc_occ_loss_bwd', ptr(p), ptr(t), ptr(g.contiguous()), p.numel(), ptr(gp), stream())
        return gp, None


class ScatterAddRows(torch.autograd.Function):
    """vol[sel] += y in place (DenseHead.py:80-81 + AdaptiveSparseHead.py:77).  ``vol`` is any contiguous [..., C]
    tensor whose leading dims flatten to the voxel index; pass the tensor itself, not a view of it (in-place on a view
    makes autograd copy the whole volume several times in the backward)."""

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


# ----------------------------------------------------------------------------------------------
# the depth-distribution producer in front of the path (csrc/sgc_depth.cu)
# ----------------------------------------------------------------------------------------------

@dataclass
class DepthCL:
    """A depth distribution already in the layout the lift kernels read: ``t`` [V, h*w, D] channel-last, cropped to (h, w).
    ``DenseHead.prepare`` takes it in place of the reference's [1,V,D,H0,W0] tensor and skips its permute copy."""
    t: torch.Tensor
    h: int
    w: int


class PlaneSweep(torch.autograd.Function):
    """corr [V,D,H,W] = plane-sweep correlation of the matching features ``feat`` [V,C,H,W] with their neighbour frames
    (depth_est_fusion.py:85-126,209-232): ``nbr`` [V,K] int32, ``rt`` [V,K,12] (rows of src_proj @ inverse(ref_proj)),
    ``depth`` [D].  One fused kernel each way on a channel-last copy of the map; the warped features never exist."""

    @staticmethod
    def forward(ctx, feat, nbr, rt, depth):
        V, C, H, W = feat.shape
        K, D = nbr.shape[1], depth.numel()
        feat = feat.contiguous()
        fcl = torch.empty(V, H * W, C, device=feat.device, dtype=F32)
        call('sgc_nchw_to_nhwc', ptr(feat), V, C, H * W, ptr(fcl), stream())
        corr = torch.empty(V, D, H, W, device=feat.device, dtype=F32)
        call('sgc_plane_sweep_fwd', ptr(fcl), ptr(nbr), ptr(rt), ptr(depth), V, K, D, H, W, C, ptr(corr), stream())
        ctx.save_for_backward(fcl, nbr, rt, depth)
        ctx.dims = (V, C, H, W, K, D)
        return corr

    @staticmethod
    def backward(ctx, gcorr):
        fcl, nbr, rt, depth = ctx.saved_tensors
        V, C, H, W, K, D = ctx.dims
        gcl = torch.empty_like(fcl)
        call('sgc_plane_sweep_bwd', ptr(fcl), ptr(nbr), ptr(rt), ptr(depth), ptr(gcorr.contiguous()), V, K, D, H, W, C, ptr(gcl),
             stream())
        gfeat = torch.empty(V, C, H, W, device=fcl.device, dtype=F32)
        call('sgc_nhwc_to_nchw', ptr(gcl), V, C, H * W, ptr(gfeat), stream())
        return gfeat, None, None, None


class DepthPyramid(torch.autograd.Function):
    """(prob [V,D,H,W], cl_0, cl_1, cl_2) from the depth logits [V,D,H,W]: softmax over D (depth_est_fusion.py:241), the
    nearest x1/2 and x1/4 levels of SGCDet.build_volume (SGCDet.py:83-85) and, per level, the channel-last crop
    [V, h*w, D] the lift kernels read -- one kernel each way.  ``crops`` = ((h, w) of the full, half and quarter level)."""

    @staticmethod
    def forward(ctx, logits, crops):
        V, D, H, W = logits.shape
        logits = logits.contiguous()
        prob = torch.empty_like(logits)
        cl = [torch.empty(V, h * w, D, device=logits.device, dtype=F32) for h, w in crops]
        (h0, w0), (h1, w1), (h2, w2) = crops
        call('sgc_depth_pyramid_fwd', ptr(logits), V, D, H, W, ptr(prob), ptr(cl[0]), h0, w0, ptr(cl[1]), h1, w1, ptr(cl[2]), h2, w2,
             stream())
        ctx.save_for_backward(logits)
        ctx.crops = crops
        ctx.set_materialize_grads(False)
        return (prob, *cl)

    @staticmethod
    def backward(ctx, gprob, g0, g1, g2):
        (logits,) = ctx.saved_tensors
        V, D, H, W = logits.shape
        (h0, w0), (h1, w1), (h2, w2) = ctx.crops
        c = lambda t: None if t is None else t.contiguous()
        gprob, g0, g1, g2 = c(gprob), c(g0), c(g1), c(g2)
        gl = torch.empty_like(logits)
        call('sgc_depth_pyramid_bwd', ptr(logits), V, D, H, W, ptr(gprob), ptr(g0), h0, w0, ptr(g1), h1, w1, ptr(g2), h2, w2, ptr(gl),
             stream())
        return gl, None
