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
HEADS_WGRAD_FP32 = _os.environ.get('SGC_HEADS_WGRAD_FP32', '0') != '0'  # per-head K/V weight grads as plain fp32 bmm
SMALL_ROWS = int(_os.environ.get('SGC_SMALL_ROWS', '0'))  # voxel-count GEMMs with at most this many rows stay plain fp32
# voxel-count GEMMs of the encoder layer on the own tcgen05 kernel (csrc/sgc_rows_gemm_tc.cu) instead of the library's
# bf16 GEMM on bf16x3 operand images
ROWS_TC = _os.environ.get('SGC_ROWS_TC', '1') != '0'
ROWS_WGRAD_TC = _os.environ.get('SGC_ROWS_WGRAD_TC', '1') != '0'  # ... and their weight gradients (sgc_rows_wgrad_tc)
# all weight gradients of a layer as ONE grouped launch at the end of its backward (sgc_rows_wgrad_group_tc); 0 = one launch
# per Linear layer as in round 1
WGRAD_GROUP = _os.environ.get('SGC_WGRAD_GROUP', '1') != '0'
# backward of the lift as a gather over pixel tiles (csrc/sgc_lift_tiles.cu) instead of the scatter kernel (REDs into a
# zero-filled grad_vg).  Parity-green, but its first version is latency-bound per tile (830 us vs 195 + 58 us of fill at the
# finest ScanNet level): off by default
LIFT_TILES = _os.environ.get('SGC_LIFT_TILES', '0') != '0'
# output_proj and the query in-projection are two back-to-back Linear layers: the chain evaluates their product
# (mean -> qv in one GEMM, W_q W_out prepared per step) and the intermediate g, needed only by the weight gradients, is
# produced off the chain on the weight-gradient stream; likewise gqv -> gmean in the backward
FUSE_QO = _os.environ.get('SGC_FUSE_QO', '0') != '0'  # measured neutral (554 vs 550-572 volumes/s): off
TOPK_MC_MIN = int(_os.environ.get('SGC_TOPK_MC_MIN', '32768'))  # levels with more voxels use the many-CTA top-k
TOPK_GRID = _os.environ.get('SGC_TOPK_GRID', '1') != '0'   # one-launch grid top-k (round 2); 0 = the round-1 kernels
_TOPK_SCRATCH = {}
ROWS_NCTA = int(_os.environ.get('SGC_ROWS_NCTA', '0'))  # output columns per CTA of that kernel (0 = its own heuristic)


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


def mm_nt(a: torch.Tensor, w: torch.Tensor = None, ws: torch.Tensor = None) -> torch.Tensor:
    """a [R,K] @ w[N,K]^T -> [R,N] fp32, tensor cores with bf16 hi/lo split (hi*hi + lo*hi + hi*lo).
    ``ws`` = pre-split weight ``split_cols(w, 1)`` (computed once per step by ``LevelWeights``)."""
    if a.shape[0] <= SMALL_ROWS and w is not None:
        return a @ w.t()  # tiny problem: the plain fp32 GEMM is launch-bound either way, skip the two splits
    if ws is None:
        ws = split_cols(w, 1)
    return torch.mm(split_cols(a, 0), ws.t(), out_dtype=F32)


def rows_linear(x: torch.Tensor, wpack: torch.Tensor, N: int, bias: Optional[torch.Tensor] = None, n_cta: int = 0):
    """y [R,N] = x [R,K] @ W^T (+ bias) with ``wpack`` = the packed [N,K] weight (``sgc_rows_gemm_tc``)."""
    R, K = x.shape
    y = torch.empty(R, N, device=x.device, dtype=F32)
    if n_cta == 0 and ROWS_NCTA and N % ROWS_NCTA == 0:
        n_cta = ROWS_NCTA
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
    """Every bf16x3 split / tcgen05 slab image of one level's weights (both orientations), produced once per step by a
    single launch -- normally on a side stream, off the critical path of the level.  Constants for the autograd
    Functions below (weight gradients are formed from the fp32 activations, not from these)."""

    def __init__(self, wcat, w_out, in_w, wo, w1, w2, num_heads: int = NUM_HEADS, images: bool = True, b_out=None,
                 in_b=None):
        """``images=False``: skip the bf16x3 images of the layer weights (operands of the library-GEMM path) wherever
        the packed operands of the own voxel-count GEMM kernel replace them."""
        with torch.no_grad():
            C = w_out.shape[0]
            dh = C // num_heads
            scale = 1.0 / math.sqrt(dh)
            wq, wk, wv = in_w[:C], in_w[C:2 * C], in_w[2 * C:]
            j = _WeightJobs(w_out.device)
            self.wcat = j.split_cols(wcat, 1)          # [N,3C]   x @ Wcat^T
            ok = wcat.shape[1] % 32 == 0 and wcat.shape[0] % 32 == 0
            self.wpack = j.pack(wcat) if ok else None
            self.wpack_t = j.pack(wcat.t()) if ok else None
            self.wcat_t = j.split_cols(wcat.t(), 1)    # [C,3N]   g @ Wcat
            self.rows_tc = ROWS_TC and C % 32 == 0 and w1.shape[0] % 32 == 0
            self.heads_tc = self.rows_tc and dh % 32 == 0
            self.fuse_qo = self.rows_tc and FUSE_QO and b_out is not None and in_b is not None
            if self.fuse_qo:
                wqo = torch.mm(wq, w_out)                        # qv = mean @ (W_q W_out)^T + (W_q b_out + b_q)
                self.bqo = torch.addmv(in_b[:C], wq, b_out)
                self.p_wqo, self.p_wqo_t = j.pack(wqo), j.pack(wqo.t())
            if self.rows_tc:
                # packed operands of sgc_rows_gemm_tc: p_x for y = a @ x^T, p_x_t for the data gradient g @ x
                self.p_w_out, self.p_w_out_t = j.pack(w_out), j.pack(w_out.t())
                self.p_wq, self.p_wq_t = j.pack(wq), j.pack(wq.t())
                self.p_wo, self.p_wo_t = j.pack(wo), j.pack(wo.t())
                self.p_w1, self.p_w1_t = j.pack(w1), j.pack(w1.t())
                self.p_w2, self.p_w2_t = j.pack(w2), j.pack(w2.t())
            # heads narrower than a 32-column k-slab (dh = 16 at C = 128): the per-head products run on the same kernel with
            # zero-extended weights (rows_heads_in_exp / rows_heads_out_exp): 8x redundant MMA work on tiny matrices instead
            # of bf16x3 operand images (315 MB each at Q = 51 200) + library GEMMs
            self.heads_exp = self.rows_tc and not self.heads_tc and dh % 8 == 0 and _os.environ.get('SGC_HEADS_EXP', '1') != '0'
            if self.heads_exp:
                hm = _head_mask(w_out.device, C, num_heads)                                          # [H, C]
                self.p_wk_in = j.pack(((wk.t() * scale).unsqueeze(0) * hm.unsqueeze(1)).reshape(num_heads * C, C))
                self.p_wv_in = j.pack((wv.t().unsqueeze(0) * hm.unsqueeze(1)).reshape(num_heads * C, C))
                self.p_wv_out = j.pack((wv.unsqueeze(1) * hm.t().unsqueeze(2)).reshape(C, num_heads * C))
                self.p_wk_out = j.pack(((wk * scale).unsqueeze(1) * hm.t().unsqueeze(2)).reshape(C, num_heads * C))
            if self.heads_tc:
                self.p_wk = j.pack(wk, scale)                          # gqv[:, h] = gqt[h] @ (scale Wk_h)^T
                self.p_wv = j.pack(wv)                                 # o[:, h]   = t[h] @ Wv_h^T
                self.p_wk_ht = j.pack_heads_t(wk, num_heads, scale)    # qt[h]     = qv_h @ (scale Wk_h)
                self.p_wv_ht = j.pack_heads_t(wv, num_heads)           # gt[h]     = go_h @ Wv_h
            for k in ('w_out', 'w_out_t', 'wq', 'wq_t', 'wo', 'wo_t', 'w1', 'w1_t', 'w2', 'w2_t', 'wk_rows', 'wk_cols',
                      'wv_rows', 'wv_cols'):
                setattr(self, k, None)
            if images or not self.rows_tc:
                self.w_out, self.w_out_t = j.split_cols(w_out, 1), j.split_cols(w_out.t(), 1)
                self.wq, self.wq_t = j.split_cols(wq, 1), j.split_cols(wq.t(), 1)
                self.wo, self.wo_t = j.split_cols(wo, 1), j.split_cols(wo.t(), 1)
                self.w1, self.w1_t = j.split_cols(w1, 1), j.split_cols(w1.t(), 1)
                self.w2, self.w2_t = j.split_cols(w2, 1), j.split_cols(w2.t(), 1)
            if images or not (self.heads_tc or self.heads_exp):
                self.wk_rows = j.split_rows(wk, dh, 1, scale)  # [8,3dh,C]  qv_h @ (scale Wk_h)
                self.wk_cols = j.split_cols(wk, 1, scale)      # [C,3C]     gqt[h] @ (scale Wk_h)^T
                self.wv_rows = j.split_rows(wv, dh, 1)     # [8,3dh,C]  go_h @ Wv_h
                self.wv_cols = j.split_cols(wv, 1)         # [C,3C]     t[h] @ Wv_h^T
            j.launch()

    def record_stream(self, s):
        for t in self.__dict__.values():
            if isinstance(t, torch.Tensor):
                t.record_stream(s)


def mm_tn(g: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """g[Q,N]^T @ x[Q,K] -> [N,K] fp32 (reduction over the rows)."""
    Q = g.shape[0]
    if Q <= SMALL_ROWS:
        return g.t() @ x
    return torch.mm(split_rows(g, Q, 0)[0].t(), split_rows(x, Q, 1)[0], out_dtype=F32)


def split_rows_colsum(g: torch.Tensor, pattern: int = 0):
    """One pass over g [Q,N]: (rows-split [3Q,N] bf16, column sums [N])."""
    g = g.contiguous()
    R, C = g.shape
    dev = g.device
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)
    ctr = _COUNTERS.get(key)
    if ctr is None:
        ctr = _COUNTERS[key] = torch.zeros(1, device=dev, dtype=torch.int32)
    out = torch.empty(3 * R, C, device=dev, dtype=BF16)
    sums = torch.empty(C, device=dev, dtype=F32)
    scratch = torch.empty(_lib.load().sgc_colsum_scratch_floats(R, C), device=dev, dtype=F32)
    call('sgc_split_rows_colsum', ptr(g), R, C, pattern, ptr(out), ptr(sums), ptr(scratch), ptr(ctr), stream())
    return out, sums


def linear_grads(g: torch.Tensor, x: torch.Tensor):
    """(gW, gb) of y = x W^T + b given g = dL/dy: gW = g^T x [N,K], gb = colsum(g); one fused pass over g."""
    Q = g.shape[0]
    if _os.environ.get('SGC_FUSED_COLSUM', '1') == '0':
        return mm_tn(g, x), colsum(g)
    gs, gb = split_rows_colsum(g, 0)
    return torch.mm(gs.t(), split_rows(x, Q, 1)[0], out_dtype=F32), gb


_COUNTERS = {}
_GRAD_STREAMS = {}


class OnStream(torch.autograd.Function):
    """Identity on parameters, applied under ``with torch.cuda.stream(wstream)``: the aliases' backward node (and the
    AccumulateGrad behind it) belong to ``wstream``, so a backward that PRODUCES a weight gradient on ``wstream`` can
    hand it to autograd without making the calling stream wait for it (autograd orders consumers after the
    producing stream, and joins every leaf stream when the backward pass ends)."""

    @staticmethod
    def forward(ctx, *ts):
        return tuple(t.view_as(t) for t in ts)

    @staticmethod
    def backward(ctx, *gs):
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
    def forward(ctx, feat: torch.Tensor, h: int, w: int, wcat: torch.Tensor, lw=None, big_stream=None):
        # feat is [V,C,H0,W0] or the reference's [1,V,C,H0,W0]; taking the 5-D leaf directly keeps autograd from
        # materialising a zero-filled copy for the select() view on the way back
        ctx.feat_shape = tuple(feat.shape)
        feat = feat.reshape(-1, *feat.shape[-3:])
        V, C, H0, W0 = feat.shape
        S = h * w
        N = wcat.shape[0]
        ctx.dims = (V, C, H0, W0, h, w)
        ctx.lw = lw
        ctx.big = big_stream   # optional dedicated stream for the two gradient kernels (see plugin._big_backward_streams)
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
            cur = torch.cuda.current_stream(gvg.device)
            big = ctx.big if ctx.big is not None and ctx.big != cur else None
            if big is not None:
                big.wait_stream(cur)
                for t_ in (gvg, feat):
                    t_.record_stream(big)
            with torch.cuda.stream(big if big is not None else cur):
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
            if big is not None:
                cur.wait_stream(big)
                for t_ in (gfeat, gw):
                    if t_ is not None:
                        t_.record_stream(cur)
            return gfeat, None, None, gw, None, None
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
        return gfeat, None, None, gw, None, None


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
# cross-view fusion
# ----------------------------------------------------------------------------------------------

def _heads_cols(x: torch.Tensor, pattern: int, heads: int = NUM_HEADS) -> torch.Tensor:
    """[Q, heads*dh] -> [heads, Q, 3dh] (strided view): per-head operand whose reduction axis is dh."""
    Q, C = x.shape
    dh = C // heads
    return split_cols(x.reshape(Q * heads, dh), pattern).view(Q, heads, 3 * dh).transpose(0, 1)


def _heads_rows_t(x: torch.Tensor, pattern: int, heads: int = NUM_HEADS) -> torch.Tensor:
    """[Q, heads*dh] -> [heads, dh, 3Q] (strided view): per-head transposed operand, reduction over Q."""
    Q, C = x.shape
    dh = C // heads
    return split_rows(x, Q, pattern)[0].view(3 * Q, heads, dh).permute(1, 2, 0)


class CrossView(torch.autograd.Function):
    """DCA:815-837: masked mean over views -> output_proj -> 8-head attention pooling over views.

    The dense projections are voxel-count GEMMs (static shapes; library bf16 GEMMs on bf16x3-split operands,
    fp32 accumulate) around the two cross-view kernels; the backward is written out by hand so that no
    pair-capacity-sized tensor is ever touched outside the kernels.
    """

    @staticmethod
    def forward(ctx, slots, pl: PairList, w_out, b_out, in_w, in_b, wo, bo, lw=None, wstream=None):
        ctx.wstream = wstream
        Q, V = pl.Q, pl.V
        C = slots.shape[1]
        H = NUM_HEADS
        if lw is None:
            lw = LevelWeights(w_out, w_out, in_w, wo, w_out, w_out)
        ctx.lw = lw
        dh = C // H
        scale = 1.0 / math.sqrt(dh)
        dev = slots.device
        mean = torch.empty(Q, C, device=dev, dtype=F32)
        call('sgc_crossview_mean_fwd', ptr(slots), ptr(pl.pair_index), V, Q, C, ptr(mean), stream())
        bq, bv = in_b[:C], in_b[2 * C:]
        wq, wk, wv = in_w[:C], in_w[C:2 * C] * scale, in_w[2 * C:]
        small = Q <= SMALL_ROWS
        g = mm_nt(mean, w_out, lw.w_out) + b_out
        qv = mm_nt(g, wq, lw.wq) + bq
        # qt[h] = qv_h @ (scale * Wk_h)   [8,Q,dh] x [8,dh,C]
        if small:
            qt = torch.bmm(qv.view(Q, H, dh).transpose(0, 1), wk.view(H, dh, C))
        else:
            qt = torch.bmm(_heads_cols(qv, 0), lw.wk_rows, out_dtype=F32)
        t = torch.empty(H, Q, C, device=dev, dtype=F32)
        alpha = torch.empty(pl.cap, H, device=dev, dtype=F32)
        call('sgc_crossview_attn_fwd', ptr(qt), ptr(slots), ptr(pl.pair_index), V, Q, C, ptr(t), ptr(alpha), stream())
        # o[h] = t[h] @ Wv_h^T   [8,Q,C] x [8,C,dh]
        if small:
            o = torch.bmm(t, wv.view(H, dh, C).transpose(1, 2))
        else:
            o = torch.bmm(split_cols(t.view(H * Q, C), 0).view(H, Q, 3 * C),
                          lw.wv_cols.view(H, dh, 3 * C).transpose(1, 2), out_dtype=F32)
        o2 = o.transpose(0, 1).reshape(Q, C) + bv
        has = (pl.count > 0).to(F32).unsqueeze(1)
        out = (mm_nt(o2, wo, lw.wo) + bo) * has
        ctx.save_for_backward(slots, mean, g, qv, qt, t, alpha, o2, has, w_out, in_w, wo)
        ctx.pl = pl
        return out

    @staticmethod
    def backward(ctx, gout):
        slots, mean, g, qv, qt, t, alpha, o2, has, w_out, in_w, wo = ctx.saved_tensors
        pl = ctx.pl
        Q, V = pl.Q, pl.V
        C = slots.shape[1]
        H = NUM_HEADS
        dh = C // H
        scale = 1.0 / math.sqrt(dh)
        dev = slots.device
        lw = ctx.lw
        wq, wk, wv = in_w[:C], in_w[C:2 * C] * scale, in_w[2 * C:]
        small = Q <= SMALL_ROWS
        side = _Side(dev, ctx.wstream)
        gout = gout * has
        g_wo, g_bo = side.run(lambda: linear_grads(gout, o2), gout, o2)
        go2 = mm_nt(gout, wo.t(), lw.wo_t)

        # gt[h] = go_h @ Wv_h   [8,Q,dh] x [8,dh,C]
        go_h = go2.view(Q, H, dh).transpose(0, 1)
        if small:
            gt = torch.bmm(go_h, wv.view(H, dh, C))
        else:
            gt = torch.bmm(_heads_cols(go2, 0), lw.wv_rows, out_dtype=F32)
        # g_wv[h] = go_h^T @ t[h]   [8,dh,Q] x [8,Q,C]
        if small:
            g_wv = torch.bmm(go_h.transpose(1, 2), t).reshape(C, C)
            g_bv = colsum(go2)
        elif HEADS_WGRAD_FP32:
            g_wv, g_bv = side.run(lambda: (torch.bmm(go_h.transpose(1, 2), t).reshape(C, C), colsum(go2)), go2, t)
        else:
            def _wv():
                gs, gb = split_rows_colsum(go2, 0)  # [3Q,C] rows-split of go2 + the bias gradient of the value proj
                a = gs.view(3 * Q, H, dh).permute(1, 2, 0)
                return torch.bmm(a, split_rows(t.view(H * Q, C), Q, 1), out_dtype=F32).reshape(C, C), gb
            g_wv, g_bv = side.run(_wv, go2, t)
        gscore = torch.empty(pl.cap, H, device=dev, dtype=F32)
        gqt = torch.empty(H, Q, C, device=dev, dtype=F32)
        call('sgc_crossview_attn_bwd_qt', ptr(slots), ptr(alpha), ptr(pl.pair_index), V, Q, C, ptr(gt), ptr(gscore),
             ptr(gqt), stream())
        # gqv[h] = gqt[h] @ (scale*Wk_h)^T   [8,Q,C] x [8,C,dh]
        if small:
            gqv_h = torch.bmm(gqt, wk.view(H, dh, C).transpose(1, 2))
        else:
            gqv_h = torch.bmm(split_cols(gqt.view(H * Q, C), 0).view(H, Q, 3 * C),
                              lw.wk_cols.view(H, dh, 3 * C).transpose(1, 2), out_dtype=F32)
        gqv = gqv_h.transpose(0, 1).reshape(Q, C)
        # g_wk[h] = scale * qv_h^T @ gqt[h]   [8,dh,Q] x [8,Q,C]
        if small:
            g_wk = torch.bmm(qv.view(Q, H, dh).permute(1, 2, 0), gqt).reshape(C, C) * scale
        elif HEADS_WGRAD_FP32:
            g_wk = side.run(lambda: torch.bmm(qv.view(Q, H, dh).permute(1, 2, 0), gqt).reshape(C, C) * scale, qv, gqt)
        else:
            g_wk = side.run(lambda: torch.bmm(_heads_rows_t(qv, 0), split_rows(gqt.view(H * Q, C), Q, 1),
                                              out_dtype=F32).reshape(C, C) * scale, qv, gqt)
        g_wq, g_bq = side.run(lambda: linear_grads(gqv, g), gqv, g)
        gg = mm_nt(gqv, wq.t(), lw.wq_t)
        g_wout, g_bout = side.run(lambda: linear_grads(gg, mean), gg, mean)
        gmean = mm_nt(gg, w_out.t(), lw.w_out_t)
        gslots = torch.empty_like(slots)
        call('sgc_crossview_attn_bwd_slots', ptr(qt), ptr(alpha), ptr(gscore), ptr(pl.pair_index), V, Q, C, ptr(gt),
             ptr(gmean), ptr(gslots), stream())
        side.join()
        g_in_w, g_in_b = side.run(lambda: (torch.cat([g_wq, g_wk, g_wv], dim=0),
                                           torch.cat([g_bq, torch.zeros_like(g_bq), g_bv], dim=0)))
        side.join()
        return gslots, None, g_wout, g_bout, g_in_w, g_in_b, g_wo, g_bo, None, None


class Linear3(torch.autograd.Function):
    """y = x W^T + b on the tensor cores with bf16x3-split operands (used for the FFN, encoder.py:335-338)."""

    @staticmethod
    def forward(ctx, x, w, b, ws=None, ws_t=None, wstream=None):
        ctx.save_for_backward(x, w)
        ctx.ws_t, ctx.wstream = ws_t, wstream
        return mm_nt(x, w, ws) + b

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gy = gy.contiguous()
        side = _Side(gy.device, ctx.wstream)
        gw, gb = side.run(lambda: linear_grads(gy, x), gy, x)
        gx = mm_nt(gy, w.t(), ctx.ws_t)
        side.join()
        return gx, gw, gb, None, None, None


class LayerNormRows(torch.autograd.Function):
    """nn.LayerNorm over voxel rows [R,C] (the norms of VoxFormerLayer, encoder.py:262-340): torch's forward kernel,
    own backward (``sgc_layernorm_bwd``: one pass for gx, the gamma/beta reduction finishes on the weight stream)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps: float, wstream=None):
        x = x.contiguous()
        y, mean, rstd = torch.native_layer_norm(x, (x.shape[1],), gamma, beta, eps)
        ctx.save_for_backward(x, mean, rstd, gamma)
        ctx.wstream = wstream
        return y

    @staticmethod
    def backward(ctx, gy):
        x, mean, rstd, gamma = ctx.saved_tensors
        R, C = x.shape
        gy = gy.contiguous()
        gx = torch.empty_like(x)
        partial = torch.empty(_lib.load().sgc_layernorm_bwd_scratch_floats(R, C), device=x.device, dtype=F32)
        call('sgc_layernorm_bwd', ptr(x), ptr(gy), ptr(mean), ptr(rstd), ptr(gamma), R, C, ptr(gx), ptr(partial), stream())
        side = _Side(x.device, ctx.wstream)

        def _params():
            gg, gb = torch.empty(C, device=x.device, dtype=F32), torch.empty(C, device=x.device, dtype=F32)
            call('sgc_layernorm_bwd_params', ptr(partial), R, C, ptr(gg), ptr(gb), stream())
            return gg, gb
        gg, gb = side.run(_params, partial)
        side.join()
        return gx, gg, gb, None, None


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
    FFN (+identity) -> LayerNorm, as ONE fixed sequence of launches: the dense GEMMs alternate with fused row kernels
    (``sgc_rowop_fwd/bwd``) that carry bias, ReLU, dropout mask, row mask, residual, LayerNorm and the bf16x3 operand
    image of the next GEMM.  The backward is written out by hand; weight / bias gradients are produced on ``wstream`` =
    (stream of the attention-block parameters, stream of the FFN / norm parameters), see ``OnStream``.

    GEMMs: tensor cores with bf16x3 operands and fp32 accumulation; levels with at most ``SMALL_ROWS`` voxels use plain
    fp32 GEMMs instead (launch-bound either way, and the library's fp32 kernels need no thread-block cluster, so they
    start at once next to the persistent projection kernels of the finer levels).

    ``masks`` = (mask_attn, mask_ffn1, mask_ffn2) uint8 keep-masks or None (eval / p = 0), ``drops`` the matching p."""

    @staticmethod
    def forward(ctx, slots, pl: PairList, w_out, b_out, in_w, in_b, wo, bo, w1, b1, w2, b2, g1, be1, g2, be2,
                lw, wstream, eps1, eps2, masks, drops, coll=None):
        """``coll`` (view sharding, ``parallel.ViewShardExchange``): ``slots`` / ``pl`` hold only the views this rank owns;
        every statistic over views becomes (local partial, written straight into peer memory) -> one all-reduce launch ->
        (local finish).  Exchanged per level: sums + counts [Q,C+1], score maxima [Q,8], the per-head value products of the
        partial softmax sums together with the normalisers [Q,C+8] (the per-head projection is applied to the PARTIAL sums:
        8x fewer bytes on the links than the [8,Q,C] sums themselves), and in the backward the normaliser dot [Q,8] and the
        query gradient after its per-head projection [Q,C]."""
        Q, V = pl.Q, pl.V
        C = slots.shape[1]
        H = NUM_HEADS
        dh = C // H
        Fh = w1.shape[0]
        dev = slots.device
        small = Q <= SMALL_ROWS
        tc = (not small) and getattr(lw, 'rows_tc', False)      # own tcgen05 GEMM for the plain Linear layers
        htc = tc and getattr(lw, 'heads_tc', False)             # ... and for the per-head key / value projections
        hexp = tc and getattr(lw, 'heads_exp', False)           # ... of heads narrower than a k-slab (zero-extended weights)
        sp = not small and not tc                               # bf16x3 operand images for the library GEMMs
        hsp = not small and not (htc or hexp)
        scale = 1.0 / math.sqrt(dh)
        bq, bv = in_b[:C], in_b[2 * C:]
        wq, wk, wv = in_w[:C], in_w[C:2 * C], in_w[2 * C:]
        m0, m1, m2 = masks if masks is not None else (None, None, None)
        s0, s1, s2 = (1.0 / (1.0 - p) if m is not None else 1.0 for m, p in zip((m0, m1, m2), drops))

        def lin(a, a_s, w, ws, pk, bias=None):  # a @ w^T (+ bias, own kernel only)
            if tc:
                return rows_linear(a, pk, w.shape[0], bias)
            return a @ w.t() if small else torch.mm(a_s, ws.t(), out_dtype=F32)

        if coll is not None and not (htc or hexp):
            raise RuntimeError('sgcdet_b200: view sharding needs the tensor-core layer path (embed_dims 128 or 256, > SMALL_ROWS rows)')
        cnt = pl.count
        if coll is None:
            mean = torch.empty(Q, C, device=dev, dtype=F32)
            mean_s = torch.empty(Q, 3 * C, device=dev, dtype=BF16) if sp else None
            call('sgc_crossview_mean_fwd_split', ptr(slots), ptr(pl.pair_index), V, Q, C, ptr(mean), ptr(mean_s), stream())
        else:
            # exchange 1: sums over the local views + local view counts -> mean over ALL views, global counts
            call('sgc_crossview_sum_fwd', ptr(slots), ptr(pl.pair_index), V, Q, C, ptr(coll.view((Q, C))), stream())
            coll.view((Q,), Q * C).copy_(pl.count)
            red = coll.reduce(Q * C + Q, 'sum')
            mean, cnt = rows_headscale(red[:Q * C].view(Q, C), red[Q * C:], 1, 1.0, want_count=True)
            mean_s = None
        fqo = htc and getattr(lw, 'fuse_qo', False) and coll is None
        if fqo:  # one GEMM on the chain; g (an input of the weight gradients only) is produced beside it
            qv, qv_hs = rows_linear(mean, lw.p_wqo, C, lw.bqo), None
            g = None
            if slots.requires_grad or w_out.requires_grad or in_w.requires_grad:
                cur = torch.cuda.current_stream(dev)
                gside = wstream[0] if wstream is not None and wstream[0] != cur else None
                if gside is not None:
                    gside.wait_stream(cur)
                    mean.record_stream(gside)
                with torch.cuda.stream(gside if gside is not None else cur):
                    g = rows_linear(mean, lw.p_w_out, C, b_out)
            else:
                g = mean.new_empty(0)
        elif tc:   # biases ride in the GEMM epilogue, no row kernel in between
            g = lin(mean, None, w_out, None, lw.p_w_out, b_out)
            if htc or hexp:
                qv, qv_hs = lin(g, None, wq, None, lw.p_wq, bq), None
            else:
                qv, qv_hs, _ = rowop_fwd(lin(g, None, wq, None, lw.p_wq), Q, C, bias=bq, split_heads=H)
        else:
            g, g_s, _ = rowop_fwd(lin(mean, mean_s, w_out, lw.w_out, None), Q, C, bias=b_out, want_split=sp)
            qv, qv_hs, _ = rowop_fwd(lin(g, g_s, wq, lw.wq, None), Q, C, bias=bq, split_heads=H, want_split=sp)
        if small:   # qt[h] = scale * qv_h @ Wk_h
            qt = torch.empty(H, Q, C, device=dev, dtype=F32)
            torch.baddbmm(qt, qv.view(Q, H, dh).transpose(0, 1), wk.view(H, dh, C), beta=0, alpha=scale, out=qt)
        elif htc:
            qt = rows_heads_in(qv, lw.p_wk_ht, C, H)
        elif hexp:
            qt = rows_heads_in_exp(qv, lw.p_wk_in, H)
        else:
            qt = torch.bmm(qv_hs.view(Q, H, 3 * dh).transpose(0, 1), lw.wk_rows, out_dtype=F32)   # [H,Q,C]
        t = torch.empty(H, Q, C, device=dev, dtype=F32)
        t_s = torch.empty(H * Q, 3 * C, device=dev, dtype=BF16) if hsp else None
        alpha = torch.empty(pl.cap, H, device=dev, dtype=F32)
        ssm = None
        if coll is not None:
            # exchange 2: score maxima; exchange 3: the partial softmax sums -- the log-sum-exp merge without a rescale pass.
            # t = this rank's UNNORMALISED partial sum_v e_v s_v, alpha = e (both finished in the backward with ssm)
            sc = torch.empty(pl.cap, H, device=dev, dtype=F32)
            call('sgc_cvs_scores', ptr(qt), ptr(slots), ptr(pl.pair_index), V, Q, C, ptr(sc), ptr(coll.view((Q, H))), stream())
            m = coll.reduce(Q * H, 'max')
            call('sgc_cvs_accum', ptr(sc), ptr(m), ptr(slots), ptr(pl.pair_index), V, Q, C, ptr(alpha),
                 ptr(coll.view((Q, H), Q * C)), ptr(t), stream())
            if htc:
                rows_heads_out(t, lw.p_wv, dh, out=coll.view((Q, C)))
            else:
                rows_heads_out_exp(t, lw.p_wv_out, out=coll.view((Q, C)))
            red = coll.reduce(Q * C + Q * H, 'sum')
            ssm = red[Q * C:].view(Q, H)
            o2, o2_s = rows_headscale(red[:Q * C].view(Q, C), ssm, H, 1e-30, bv), None
        else:
            call('sgc_crossview_attn_fwd_split', ptr(qt), ptr(slots), ptr(pl.pair_index), V, Q, C, ptr(t), ptr(alpha), ptr(t_s),
                 stream())
        if coll is not None:
            pass
        elif htc:   # o2[:, h] = t[h] @ Wv_h^T + bv_h, written straight into the [Q,C] layout
            o2, o2_s = rows_heads_out(t, lw.p_wv, dh, bv), None
        elif hexp:
            o2, o2_s = rows_heads_out_exp(t, lw.p_wv_out, bv), None
        else:
            if small:   # o[h] = t[h] @ Wv_h^T   [H,Q,dh]
                o = torch.bmm(t, wv.view(H, dh, C).transpose(1, 2))
            else:
                o = torch.bmm(t_s.view(H, Q, 3 * C), lw.wv_cols.view(H, dh, 3 * C).transpose(1, 2), out_dtype=F32)
            o2, o2_s, _ = rowop_fwd(o, Q, C, bias=bv, in_heads=H, want_split=sp)
        # rows no view sees are zeroed (DCA:819-835) straight from the per-voxel view count
        x1, x1_s, ln1 = rowop_fwd(lin(o2, o2_s, wo, lw.wo, getattr(lw, 'p_wo', None)), Q, C, bias=bo, mask=m0, mscale=s0,
                                  rowcount=cnt, ln=(g1, be1, eps1), want_split=sp)
        hdn, hdn_s, _ = rowop_fwd(lin(x1, x1_s, w1, lw.w1, getattr(lw, 'p_w1', None)), Q, Fh, bias=b1, relu=True, mask=m1,
                                  mscale=s1, want_split=sp)
        y, _, ln2 = rowop_fwd(lin(hdn, hdn_s, w2, lw.w2, getattr(lw, 'p_w2', None)), Q, C, bias=b2, mask=m2, mscale=s2,
                              residual=x1, ln=(g2, be2, eps2), want_split=False)
        ctx.save_for_backward(slots, mean, g, qv, qt, t, alpha, o2, x1, hdn, *ln1, *ln2, g1, g2,
                              w_out, in_w, wo, w1, w2)
        ctx.pl, ctx.lw, ctx.wstream = pl, lw, wstream
        ctx.coll, ctx.ssm, ctx.cnt = coll, ssm, cnt
        ctx.masks, ctx.scales = (m0, m1, m2), (s0, s1, s2)
        return y

    @staticmethod
    def backward(ctx, gy):
        (slots, mean, g, qv, qt, t, alpha, o2, x1, hdn, pre1, mean1, rstd1, pre2, mean2, rstd2, g1, g2,
         w_out, in_w, wo, w1, w2) = ctx.saved_tensors
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
        small = Q <= SMALL_ROWS
        tc = (not small) and getattr(lw, 'rows_tc', False)
        htc = tc and getattr(lw, 'heads_tc', False)
        hexp = tc and getattr(lw, 'heads_exp', False)
        sp = not small and not tc
        hsp = not small and not (htc or hexp)
        fp32_heads = small or HEADS_WGRAD_FP32
        wq, wk, wv = in_w[:C], in_w[C:2 * C], in_w[2 * C:]
        ws_attn, ws_ffn = ctx.wstream if ctx.wstream is not None else (None, None)
        side = _Side(dev, ws_attn)     # attention-block parameters
        side_f = _Side(dev, ws_ffn)    # FFN + norms
        gy = gy.contiguous()

        def lin_t(a, a_s, w, ws_t, pk_t):  # a @ w
            if tc:
                return rows_linear(a, pk_t, w.shape[1])
            return a @ w if small else torch.mm(a_s, ws_t.t(), out_dtype=F32)

        wtc = tc and ROWS_WGRAD_TC and C % 128 == 0 and Fh % 128 == 0   # own kernel for the weight gradients too
        # all seven weight-gradient products of the layer as ONE grouped launch at the end of this backward (also the per-head
        # key / value products of the 16-wide heads, which the per-layer launches leave to the library)
        grouped = wtc and WGRAD_GROUP and not getattr(lw, 'fuse_qo', False) and dh % 16 == 0
        if coll is not None and not grouped:
            raise RuntimeError('sgcdet_b200: view sharding needs the grouped weight-gradient launch (SGC_WGRAD_GROUP=1)')
        hwtc = wtc and htc and not grouped
        lgrads = linear_grads_tc if wtc else linear_grads
        if True:
            # norm 2 + dropout of the second FFN layer; gpre2 also flows into the identity branch
            gf, gf_s, gpre2, part2 = rowop_bwd(gy, Q, C, ln=(pre2, mean2, rstd2, g2), mask=m2, mscale=s2, want_gpre=True,
                                               want_split=sp)
            g_g2, g_be2 = side_f.run(lambda: _ln_params(part2, Q, C), part2)
            if not grouped:
                g_w2, g_b2 = side_f.run(lambda: lgrads(gf, hdn), gf, hdn)
            ghdn = lin_t(gf, gf_s, w2, lw.w2_t, getattr(lw, 'p_w2_t', None))                            # [Q,F]
            # hdn = relu(.)*mask1*s1, so the ReLU gate and the dropout mask together are (hdn > 0)
            gh, gh_s, _, _ = rowop_bwd(ghdn, Q, Fh, gate=hdn, gscale=s1, want_split=sp)
            if not grouped:
                g_w1, g_b1 = side_f.run(lambda: lgrads(gh, x1), gh, x1)
            gx1_raw = lin_t(gh, gh_s, w1, lw.w1_t, getattr(lw, 'p_w1_t', None))                         # [Q,C]
            gout, gout_s, _, part1 = rowop_bwd(gx1_raw, Q, C, g2=gpre2, ln=(pre1, mean1, rstd1, g1), mask=m0, mscale=s0,
                                               rowcount=cnt, want_split=sp)
            g_g1, g_be1 = side_f.run(lambda: _ln_params(part1, Q, C), part1)
            if not grouped:
                g_wo, g_bo = side.run(lambda: lgrads(gout, o2), gout, o2)
            go2 = lin_t(gout, gout_s, wo, lw.wo_t, getattr(lw, 'p_wo_t', None))                         # [Q,C]
        if hwtc:
            # the three in-projection gradients are written straight into in_proj_weight's / in_proj_bias's gradients
            # (rows [0,C) query, [C,2C) key, [2C,3C) value; the key bias gradient is identically zero)
            def _alloc_in():
                gw_ = torch.empty(3 * C, C, device=dev, dtype=F32)
                gb_ = torch.empty(3 * C, device=dev, dtype=F32)
                gb_[C:2 * C].zero_()
                return gw_, gb_
            g_in_w, g_in_b = side.run(_alloc_in)
        if small:   # gt[h] = go_h @ Wv_h
            gt = torch.bmm(go2.view(Q, H, dh).transpose(0, 1), wv.view(H, dh, C))
        elif htc:
            gt = rows_heads_in(go2, lw.p_wv_ht, C, H)                                               # [H,Q,C]
        elif hexp:
            gt = rows_heads_in_exp(go2, lw.p_wv_in, H)
        else:
            _, go2_hs, _, _ = rowop_bwd(go2, Q, C, split_heads=H, want_gx=False)
            gt = torch.bmm(go2_hs.view(Q, H, 3 * dh).transpose(0, 1), lw.wv_rows, out_dtype=F32)    # [H,Q,C]

        def _wv():
            if hwtc:   # g_wv[h*dh + d, c] = sum_q t[h][q, c] go2[q, h*dh + d];  g_bv = column sums of go2
                return (rows_wgrad(t, go2, C, dh, Q, g_in_w[2 * C:], (dh * C, 1, C), B=H, lda=C, batch_a=Q * C, ldb=C,
                                   batch_b=dh, bias_out=g_in_b[2 * C:], bias_from=2), g_in_b[2 * C:])
            if fp32_heads:
                return torch.bmm(go2.view(Q, H, dh).permute(1, 2, 0), t).reshape(C, C), colsum(go2)
            gs, gb = split_rows_colsum(go2, 0)
            a = gs.view(3 * Q, H, dh).permute(1, 2, 0)
            return torch.bmm(a, split_rows(t.view(H * Q, C), Q, 1), out_dtype=F32).reshape(C, C), gb
        if not grouped:
            g_wv, g_bv = side.run(_wv, go2, t)
        gscore = torch.empty(pl.cap, H, device=dev, dtype=F32)
        gqt = torch.empty(H, Q, C, device=dev, dtype=F32)
        gqt_s = torch.empty(H * Q, 3 * C, device=dev, dtype=BF16) if hsp else None
        if coll is not None:
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
            if htc:
                rows_heads_out(gqt, lw.p_wk, dh, out=coll.view((Q, C)))
            else:
                rows_heads_out_exp(gqt, lw.p_wk_out, out=coll.view((Q, C)))
            gqv, gqv_s = coll.reduce(Q * C, 'sum').view(Q, C), None
            # the value-weight gradient pairs this rank's unnormalised partial t with go2 / ssm (per head): sum over ranks =
            # go_h^T (sum_ranks t / ssm)
            go2_n = rows_headscale(go2, ssm, H, 1e-30)
        else:
            call('sgc_crossview_attn_bwd_qt_split', ptr(slots), ptr(alpha), ptr(pl.pair_index), V, Q, C, ptr(gt), ptr(gscore),
                 ptr(gqt), ptr(gqt_s), stream())
        if coll is not None:
            pass
        elif htc:   # gqv[:, h] = gqt[h] @ (scale Wk_h)^T, written straight into the [Q,C] layout
            gqv, gqv_s = rows_heads_out(gqt, lw.p_wk, dh), None
        elif hexp:
            gqv, gqv_s = rows_heads_out_exp(gqt, lw.p_wk_out), None
        else:
            if small:   # gqv[h] = scale * gqt[h] @ Wk_h^T   [H,Q,dh]
                gqv_h = torch.empty(H, Q, dh, device=dev, dtype=F32)
                torch.baddbmm(gqv_h, gqt, wk.view(H, dh, C).transpose(1, 2), beta=0, alpha=scale, out=gqv_h)
            else:
                gqv_h = torch.bmm(gqt_s.view(H, Q, 3 * C), lw.wk_cols.view(H, dh, 3 * C).transpose(1, 2), out_dtype=F32)
            gqv, gqv_s, _, _ = rowop_bwd(gqv_h, Q, C, in_heads=H, want_split=sp)

        def _wk():
            if hwtc:   # g_wk[h*dh + d, c] = scale * sum_q gqt[h][q, c] qv[q, h*dh + d]
                return rows_wgrad(gqt, qv, C, dh, Q, g_in_w[C:2 * C], (dh * C, 1, C), B=H, lda=C, batch_a=Q * C, ldb=C,
                                  batch_b=dh, scale=scale)
            if fp32_heads:
                return torch.bmm(qv.view(Q, H, dh).permute(1, 2, 0), gqt).reshape(C, C) * scale
            return torch.bmm(_heads_rows_t(qv, 0), split_rows(gqt.view(H * Q, C), Q, 1), out_dtype=F32).reshape(C, C) * scale
        if grouped:
            pass
        elif hwtc:
            g_wk = side.run(_wk, qv, gqt)
            g_wq, g_bq = side.run(lambda: linear_grads_tc(gqv, g, g_in_w[:C], g_in_b[:C]), gqv, g)
        else:
            g_wk = side.run(_wk, qv, gqt)
            g_wq, g_bq = side.run(lambda: lgrads(gqv, g), gqv, g)
        if htc and getattr(lw, 'fuse_qo', False):
            # chain: gmean = gqv @ (W_q W_out) in one GEMM; gg = gqv @ W_q (an input of output_proj's weight gradient only)
            # is produced on the weight-gradient stream
            gmean = rows_linear(gqv, lw.p_wqo_t, C)
            g_wout, g_bout = side.run(lambda: lgrads(rows_linear(gqv, lw.p_wq_t, C), mean), gqv, mean)
        else:
            gg = lin_t(gqv, gqv_s, wq, lw.wq_t, getattr(lw, 'p_wq_t', None))
            gg_s = split_cols(gg, 0) if sp else None
            if not grouped:
                g_wout, g_bout = side.run(lambda: lgrads(gg, mean), gg, mean)
            gmean = lin_t(gg, gg_s, w_out, lw.w_out_t, getattr(lw, 'p_w_out_t', None))
        gslots = torch.empty_like(slots)
        if coll is not None:
            call('sgc_cvs_bwd_slots', ptr(qt), ptr(alpha), ptr(gscore), ptr(pl.pair_index), V, Q, C, ptr(gt), ptr(gmean),
                 ptr(cnt), ptr(gslots), stream())
        else:
            call('sgc_crossview_attn_bwd_slots', ptr(qt), ptr(alpha), ptr(gscore), ptr(pl.pair_index), V, Q, C, ptr(gt),
                 ptr(gmean), ptr(gslots), stream())
        if grouped:
            def _group():
                G = WgradGroup(Q, dev)
                gw_in = torch.empty(3 * C, C, device=dev, dtype=F32)
                gb_in = torch.zeros(3 * C, device=dev, dtype=F32)    # the key bias gradient is identically zero
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
        if not hwtc and not grouped:
            g_in_w, g_in_b = side.run(lambda: (torch.cat([g_wq, g_wk, g_wv], dim=0),
                                               torch.cat([g_bq, torch.zeros_like(g_bq), g_bv], dim=0)))
        side.join()
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
        call('sgc_upsample2x_occ_bwd', ptr(vol), X, Y, Z, C, ptr(w_occ), ptr(occ), ptr(gup),
             ptr(gocc.contiguous()) if gocc is not None else None, ptr(gpre), ptr(gin),
             None if detached else ptr(gw), ptr(gb), stream())
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
