"""The four functions of the reference's pybind module ``dfa3D._ext`` (csrc/pybind.cpp:42-67), same names,
argument order, keyword names and ownership rules, implemented over the sgcdet_b200 C ABI.

  * every tensor must be a contiguous CUDA tensor (WMSL:220-238) -> RuntimeError otherwise;
  * ``batch % min(batch, im2col_step) == 0`` is checked like WMSL:250-253 (the value has no numerical effect);
  * forward functions allocate and return their output (WMSL:255-256, DSL:84-85);
  * backward functions accumulate into caller-allocated, caller-zeroed grads (F3D:319-322,338-339);
  * float32 and float64 like the reference's AT_DISPATCH_FLOATING_TYPES (the SGCDet path itself is fp32; the fp64
    instantiation is a plain scalar one, csrc/dfa3d_op_f64.cu); the one-stage additions at the end are fp32 only.
"""
import torch

from sgcdet_b200._lib import call, ptr, stream


def _check(im2col_step, *tensors):
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError('implementation for device cpu not found (sgcdet_b200 is CUDA only)')
        if not t.is_contiguous():
            raise RuntimeError('tensor has to be contiguous')
    batch = tensors[0].size(0)
    step = min(batch, int(im2col_step))
    if batch % step != 0:
        raise RuntimeError(f'batch({batch}) must divide im2col_step({step})')


def _f32(*ts):
    for t in ts:
        if t.dtype != torch.float32:
            raise RuntimeError('sgcdet_b200: the one-stage DFA3D kernels are fp32 only')


def _suffix(*ts):
    """'' for float32, '_f64' for float64 operands (all of one dtype), like AT_DISPATCH_FLOATING_TYPES."""
    dt = ts[0].dtype
    if dt not in (torch.float32, torch.float64) or any(t.dtype != dt for t in ts):
        raise RuntimeError(f'sgcdet_b200 DFA3D kernels are implemented for float32 and float64 operands of one dtype (got '
                           f'{[str(t.dtype) for t in ts]})')
    return '' if dt == torch.float32 else '_f64'


def wms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                            attention_weights, depth_scores, im2col_step):
    _check(im2col_step, value, value_spatial_shapes, value_level_start_index, sampling_locations,
           attention_weights, depth_scores)
    sfx = _suffix(value, sampling_locations, attention_weights, depth_scores)
    B, S, M, Cm = value.shape
    L = value_spatial_shapes.size(0)
    Q, P = sampling_locations.size(1), sampling_locations.size(4)
    out = torch.empty(B, Q, M * Cm, device=value.device, dtype=value.dtype)
    with torch.cuda.device(value.device):
        call('dfa3d_wms_fwd' + sfx, ptr(value), ptr(value_spatial_shapes), ptr(value_level_start_index),
             ptr(sampling_locations), ptr(attention_weights), ptr(depth_scores), B, S, M, Cm, L, Q, P, ptr(out),
             stream())
    return out


def wms_deform_attn_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                             attention_weights, depth_scores, grad_output, grad_value, grad_sampling_loc,
                             grad_attn_weight, grad_depth_score, im2col_step):
    _check(im2col_step, value, value_spatial_shapes, value_level_start_index, sampling_locations,
           attention_weights, depth_scores, grad_output, grad_value, grad_sampling_loc, grad_attn_weight,
           grad_depth_score)
    sfx = _suffix(value, sampling_locations, attention_weights, depth_scores, grad_output, grad_value, grad_sampling_loc,
                  grad_attn_weight, grad_depth_score)
    B, S, M, Cm = value.shape
    L = value_spatial_shapes.size(0)
    Q, P = sampling_locations.size(1), sampling_locations.size(4)
    with torch.cuda.device(value.device):
        call('dfa3d_wms_bwd' + sfx, ptr(value), ptr(value_spatial_shapes), ptr(value_level_start_index),
             ptr(sampling_locations), ptr(attention_weights), ptr(depth_scores), ptr(grad_output), B, S, M, Cm, L, Q, P,
             ptr(grad_value), ptr(grad_sampling_loc), ptr(grad_attn_weight), ptr(grad_depth_score), stream())


def ms_depth_score_sample_forward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                                  im2col_step):
    _check(im2col_step, value, value_spatial_shapes, value_level_start_index, sampling_locations)
    sfx = _suffix(value, sampling_locations)
    B, S, M, D = value.shape
    L = value_spatial_shapes.size(0)
    Q, P = sampling_locations.size(1), sampling_locations.size(4)
    out = torch.empty(B, Q, M, L, P, 4, device=value.device, dtype=value.dtype)
    with torch.cuda.device(value.device):
        call('dfa3d_depth_score_fwd' + sfx, ptr(value), ptr(value_spatial_shapes), ptr(value_level_start_index),
             ptr(sampling_locations), B, S, M, D, L, Q, P, ptr(out), stream())
    return out


def ms_depth_score_sample_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                                   grad_output, grad_value, grad_sampling_loc, im2col_step):
    _check(im2col_step, value, value_spatial_shapes, value_level_start_index, sampling_locations, grad_output,
           grad_value, grad_sampling_loc)
    sfx = _suffix(value, sampling_locations, grad_output, grad_value, grad_sampling_loc)
    B, S, M, D = value.shape
    L = value_spatial_shapes.size(0)
    Q, P = sampling_locations.size(1), sampling_locations.size(4)
    with torch.cuda.device(value.device):
        call('dfa3d_depth_score_bwd' + sfx, ptr(value), ptr(value_spatial_shapes), ptr(value_level_start_index),
             ptr(sampling_locations), ptr(grad_output), B, S, M, D, L, Q, P, ptr(grad_value), ptr(grad_sampling_loc),
             stream())


# --- additions (not in the reference): the one-stage operator without the depth-score round trip --------

def dfa3d_fused_forward(value, value_dpt_dist, spatial_shapes_3d, level_start_index, sampling_locations,
                        attention_weights, need_depth_score=True):
    _check(1 << 30, value, value_dpt_dist, spatial_shapes_3d, level_start_index, sampling_locations, attention_weights)
    _f32(value, value_dpt_dist, sampling_locations, attention_weights)
    B, S, M, Cm = value.shape
    D = value_dpt_dist.size(3)
    L = spatial_shapes_3d.size(0)
    Q, P = sampling_locations.size(1), sampling_locations.size(4)
    out = torch.empty(B, Q, M * Cm, device=value.device, dtype=value.dtype)
    ds = torch.empty(B, Q, M, L, P, 4, device=value.device, dtype=value.dtype) if need_depth_score else None
    with torch.cuda.device(value.device):
        call('dfa3d_fused_fwd', ptr(value), ptr(value_dpt_dist), ptr(spatial_shapes_3d), ptr(level_start_index),
             ptr(sampling_locations), ptr(attention_weights), B, S, M, Cm, D, L, Q, P, ptr(out), ptr(ds), stream())
    return out, ds


def dfa3d_fused_backward(value, value_dpt_dist, spatial_shapes_3d, level_start_index, sampling_locations,
                         attention_weights, grad_output, grad_value, grad_dist, grad_sampling_loc, grad_attn_weight):
    _check(1 << 30, value, value_dpt_dist, spatial_shapes_3d, level_start_index, sampling_locations, attention_weights,
           grad_output, grad_value, grad_dist, grad_sampling_loc, grad_attn_weight)
    B, S, M, Cm = value.shape
    D = value_dpt_dist.size(3)
    L = spatial_shapes_3d.size(0)
    Q, P = sampling_locations.size(1), sampling_locations.size(4)
    with torch.cuda.device(value.device):
        call('dfa3d_fused_bwd', ptr(value), ptr(value_dpt_dist), ptr(spatial_shapes_3d), ptr(level_start_index),
             ptr(sampling_locations), ptr(attention_weights), ptr(grad_output), B, S, M, Cm, D, L, Q, P,
             ptr(grad_value), ptr(grad_dist), ptr(grad_sampling_loc), ptr(grad_attn_weight), stream())
