"""ctypes binding of include/sgcdet_b200.h.  There is NO fallback: if the CUDA library is missing or a
tensor is not a contiguous CUDA tensor the call raises."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_float, c_int, c_longlong, c_void_p
from pathlib import Path

import torch

_LIB_PATH = Path(__file__).resolve().parent / '_C' / 'libsgcdet_b200.so'
_lib = None

P = c_void_p
I = c_int
F = c_float
LL = c_longlong

# name -> argtypes (must mirror include/sgcdet_b200.h)
SIGNATURES = {
    'dfa3d_depth_score_fwd': [P, P, P, P, I, I, I, I, I, I, I, P, P],
    'dfa3d_depth_score_bwd': [P, P, P, P, P, I, I, I, I, I, I, I, P, P, P],
    'dfa3d_wms_fwd': [P, P, P, P, P, P, I, I, I, I, I, I, I, P, P],
    'dfa3d_wms_bwd': [P, P, P, P, P, P, P, I, I, I, I, I, I, I, P, P, P, P, P],
    'dfa3d_depth_score_fwd_f64': [P, P, P, P, I, I, I, I, I, I, I, P, P],
    'dfa3d_depth_score_bwd_f64': [P, P, P, P, P, I, I, I, I, I, I, I, P, P, P],
    'dfa3d_wms_fwd_f64': [P, P, P, P, P, P, I, I, I, I, I, I, I, P, P],
    'dfa3d_wms_bwd_f64': [P, P, P, P, P, P, P, I, I, I, I, I, I, I, P, P, P, P, P],
    'dfa3d_fused_fwd': [P, P, P, P, P, P, I, I, I, I, I, I, I, I, P, P, P],
    'dfa3d_fused_bwd': [P, P, P, P, P, P, P, I, I, I, I, I, I, I, I, P, P, P, P, P],
    'sgc_project_scratch_ints': [I, I],
    'sgc_project_compact': [P, P, P, I, I, F, F, F, F, F, F, F, F, F, P, P, P, P, P, P, P, P],
    'sgc_split_bf16x3': [P, LL, I, LL, I, I, P, P],
    'sgc_pack_weight_tc': [P, I, I, P, P],
    'sgc_prepare_weights': [P, I, P],
    'sgc_rowop_fwd': [P, P],
    'sgc_rowop_bwd': [P, P],
    'sgc_dropout_masks': [P, I, LL, P, P],
    'sgc_crossview_mean_fwd_split': [P, P, I, I, I, P, P, P],
    'sgc_crossview_attn_fwd_split': [P, P, P, I, I, I, P, P, P, P],
    'sgc_crossview_attn_bwd_qt_split': [P, P, P, I, I, I, P, P, P, P, P],
    'sgc_layernorm_bwd_scratch_floats': [I, I],
    'sgc_layernorm_bwd': [P, P, P, P, P, I, I, P, P, P],
    'sgc_layernorm_bwd_params': [P, I, I, P, P, P],
    'sgc_project_tc_fwd': [P, LL, LL, I, I, I, P, I, P, P],
    'sgc_project_tc_bwd_data': [P, I, I, I, P, I, P, LL, P],
    'sgc_project_tc_wgrad_scratch_floats': [I, I],
    'sgc_project_tc_set_max_ctas': [I],
    'sgc_project_tc_set_max_ctas_fwd': [I],
    'sgc_project_tc_set_tiles_per_cta': [I],
    'sgc_set_pdl': [I],
    'sgc_project_tc_wgrad': [P, P, LL, I, I, I, I, P, P, P],
    'sgc_rows_gemm_tc_auto_ncta': [I, I, I],
    'sgc_rows_gemm_tc_set_debug': [P],
    'sgc_rows_gemm_tc': [P, LL, LL, I, I, I, P, I, LL, I, P, I, I, P, LL, LL, I, P],
    'sgc_rows_gemm_tc_ex': [P, LL, LL, I, I, I, P, I, LL, I, P, I, I, P, LL, LL, I, I, I, P],
    'sgc_rows_wgrad_tc_scratch_floats': [I, I, I, I],
    'sgc_rows_wgrad_tc': [P, LL, LL, I, P, LL, LL, I, I, I, P, LL, LL, LL, F, P, I, P, P],
    'sgc_fold_wcat': [P, P, P, P, P, P, P, I, I, P, P, P],
    'sgc_unfold_wcat_grad': [P, P, I, I, P, P, P, P, P, P, P, P],
    'sgc_rows_wgrad_group_scratch_floats': [P, I, I],
    'sgc_rows_wgrad_group_tc': [P, I, I, P, P],
    'sgc_colsum_scratch_floats': [I, I],
    'sgc_colsum': [P, I, I, P, P, P, P],
    'sgc_split_rows_colsum': [P, I, I, I, P, P, P, P, P],
    'sgc_lift_fwd': [P, I, P, I, P, P, P, P, P, I, P, I, I, I, I, I, I, P, P, P],
    'sgc_lift_bwd_tiles_workspace_bytes': [I, I, I, I, I],
    'sgc_lift_bwd_tiles': [P, I, P, I, P, P, P, P, I, P, P, P, I, I, I, I, I, I, I, P, P, P, P, P, P, P],
    'sgc_lift_bwd_scratch_floats': [I, I],
    'sgc_lift_bwd': [P, I, P, I, P, P, P, P, I, P, P, P, I, I, I, I, I, I, P, P, P, P, P, P, P],
    'sgc_crossview_mean_fwd': [P, P, I, I, I, P, P],
    'sgc_crossview_attn_fwd': [P, P, P, I, I, I, P, P, P],
    'sgc_crossview_attn_bwd_qt': [P, P, P, I, I, I, P, P, P, P],
    'sgc_crossview_attn_bwd_slots': [P, P, P, P, I, I, I, P, P, P, P],
    'sgc_crossview_sum_fwd': [P, P, I, I, I, P, P],
    'sgc_cvs_scores': [P, P, P, I, I, I, P, P, P],
    'sgc_cvs_accum': [P, P, P, P, I, I, I, P, P, P, P],
    'sgc_cvs_bwd_dot': [P, P, P, P, I, I, I, P, P, P, P, P],
    'sgc_cvs_bwd_qt': [P, P, P, P, P, I, I, I, P, P, P],
    'sgc_cvs_bwd_slots': [P, P, P, P, I, I, I, P, P, P, P, P],
    'sgc_rows_headscale': [P, P, I, F, P, I, I, P, P, P],
    'sgc_upsample2x_occ_fwd': [P, I, I, I, I, P, P, P, P, P],
    'sgc_upsample2x_occ_bwd': [P, I, I, I, I, P, P, P, P, P, P, P, P, P, P],
    'sgc_upsample2x_occ_bwd_scratch_floats': [I, I, I, I],
    'sgc_upsample2x_occ_gradw': [P, I, I, I, I, P, P, P],
    'sgc_topk_select': [P, I, I, P, P, P],
    'sgc_topk_scratch_ints': [I],
    'sgc_topk_select_mc': [P, I, I, P, P, P, P],
    'sgc_topk_grid_scratch_bytes': [],
    'sgc_topk_grid_max_n': [],
    'sgc_topk_select_grid': [P, I, I, P, P, P, P],
    'sgc_occ_loss_fwd': [P, P, I, P, P],
    'sgc_occ_loss_bwd': [P, P, P, I, P, P],
    'sgc_valid_pyramid': [P, I, I, I, P, P, P, P],
    'sgc_plane_sweep_fwd': [P, P, P, P, I, I, I, I, I, I, P, P],
    'sgc_plane_sweep_bwd': [P, P, P, P, P, I, I, I, I, I, I, P, P],
    'sgc_nchw_to_nhwc': [P, I, I, I, P, P],
    'sgc_nhwc_to_nchw': [P, I, I, I, P, P],
    'sgc_depth_pyramid_fwd': [P, I, I, I, I, P, P, I, I, P, I, I, P, I, I, P],
    'sgc_depth_pyramid_bwd': [P, I, I, I, I, P, P, I, I, P, I, I, P, I, I, P, P],
    'sgc_peer_allreduce': [P, P, I, I, LL, I, F, P, I, I, P],
    'sgc_peer_copy_segments': [P, P, P, I, I, P],
    'sgc_peer_allreduce_tensors': [P, P, I, I, P, P, I, F, P],
    'sgc_peer_sig_bytes': [],
    'sgc_peer_status_offset': [],
    'sgc_peer_alloc': [LL, P, P],
    'sgc_peer_open': [P, P],
    'sgc_peer_close': [P],
    'sgc_peer_free': [P],
    'sgc_scatter_add_rows': [P, P, P, I, I, P],
    'sgc_gather_rows': [P, P, P, I, I, P],
}


class MaskJob(ctypes.Structure):
    """``sgc_mask_job`` of include/sgcdet_b200.h."""
    _fields_ = [('out', c_void_p), ('n', c_longlong), ('keep', c_float)]


class WeightJob(ctypes.Structure):
    """``sgc_weight_job`` of include/sgcdet_b200.h."""
    _fields_ = [('src', c_void_p), ('out', c_void_p), ('row_stride', c_longlong), ('col_stride', c_longlong),
                ('rows', c_int), ('cols', c_int), ('rows_per_group', c_int), ('pattern', c_int),
                ('scale', c_float), ('kind', c_int)]


MAX_WEIGHT_JOBS = 48


class WgradJob(ctypes.Structure):
    """``sgc_wgrad_job`` of include/sgcdet_b200.h."""
    _fields_ = [('a', c_void_p), ('lda', c_longlong), ('batch_a', c_longlong), ('M', c_int),
                ('b', c_void_p), ('ldb', c_longlong), ('batch_b', c_longlong), ('N', c_int),
                ('B', c_int),
                ('out', c_void_p), ('out_b', c_longlong), ('out_m', c_longlong), ('out_n', c_longlong),
                ('scale', c_float),
                ('bias_out', c_void_p), ('bias_from', c_int)]


MAX_WGRAD_JOBS = 8


class RowopFwdArgs(ctypes.Structure):
    """``sgc_rowop_fwd_args`` of include/sgcdet_b200.h."""
    _fields_ = [(n, c_void_p) for n in ('x', 'bias', 'mask', 'rowscale', 'residual', 'gamma', 'beta', 'y', 'ysplit',
                                        'pre', 'mean', 'rstd')] + \
               [('mscale', c_float), ('eps', c_float), ('R', c_int), ('N', c_int), ('relu', c_int),
                ('in_heads', c_int), ('split_heads', c_int), ('rowcount', c_void_p)]


class RowopBwdArgs(ctypes.Structure):
    """``sgc_rowop_bwd_args`` of include/sgcdet_b200.h."""
    _fields_ = [(n, c_void_p) for n in ('g', 'g2', 'pre', 'mean', 'rstd', 'gamma', 'mask', 'gate', 'rowscale',
                                        'partial', 'gpre', 'gx', 'gxsplit')] + \
               [('mscale', c_float), ('gscale', c_float), ('R', c_int), ('N', c_int), ('in_heads', c_int),
                ('split_heads', c_int), ('rowcount', c_void_p)]


def lib_path() -> Path:
    return _LIB_PATH


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise RuntimeError(
                f'{_LIB_PATH} is missing: build it with `python -m sgcdet_b200.build` '
                '(sgcdet_b200 has no CPU or PyTorch fallback)')
        # a stale library (older argument structs / signatures than this tree declares) would be called through ctypes
        # with mismatched layouts and corrupt memory silently: the build stamp must equal the digest of the sources
        from . import build as _build
        if os.environ.get('SGC_ALLOW_STALE_LIB', '0') == '0':
            stamp = _build.STAMP.read_text().strip() if _build.STAMP.exists() else None
            if stamp != _build._digest():
                raise RuntimeError(
                    f'{_LIB_PATH} was built from different sources than this tree (build.stamp mismatch): rebuild it with '
                    '`python -m sgcdet_b200.build`')
        lib = ctypes.CDLL(str(_LIB_PATH))
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = c_longlong if name in ('sgc_rows_wgrad_group_scratch_floats', 'sgc_lift_bwd_tiles_workspace_bytes',
                                                 'sgc_upsample2x_occ_bwd_scratch_floats') else c_int
        # measured settings (profiles/README.md): the forward projection leaves 16 SMs to the coarser levels' voxel chains (caps
        # of 100..148 measured: 132 best), no cap for the gradient kernels, no programmatic dependent launch (neutral under
        # graph replay), persistent CTAs walk all their tiles
        lib.sgc_project_tc_set_max_ctas(0)
        lib.sgc_project_tc_set_max_ctas_fwd(132)
        lib.sgc_set_pdl(0)
        lib.sgc_project_tc_set_tiles_per_cta(0)
        _lib = lib
    return _lib


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError('sgcdet_b200: expected a CUDA tensor (there is no CPU implementation)')
    if not t.is_contiguous():
        raise RuntimeError('sgcdet_b200: tensor has to be contiguous')
    return t.data_ptr()


def stream(device=None):
    """Raw handle of torch's current stream on ``device`` (default: the current device).  The module entry points
    (``plugin.AdaptiveSparseHead.forward``, ``DenseHead.forward``, the dropin
    ``_ext`` wrappers) make the tensors' device current around their launches, so the default is the tensors' device."""
    return torch.cuda.current_stream(device).cuda_stream


def call(name: str, *args):
    rc = getattr(load(), name)(*args)
    if rc != 0:
        raise RuntimeError(f'sgcdet_b200.{name} failed with cudaError {rc}')
