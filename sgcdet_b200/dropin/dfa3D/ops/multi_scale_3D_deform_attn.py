"""The three autograd Functions of ``dfa3D/ops/multi_scale_3D_deform_attn.py`` (same names and ``apply``
signatures).  Gradient stitching follows the plugin's corrected copy
(``multi_scale_3ddeformable_attn_function.py:303-351``), not the upstream file, whose backward drops the uv
gradients (``multi_scale_3D_deform_attn.py:201``) and returns 8 grads for 7 inputs (``:220-221``)."""
import torch
from torch.amp import custom_bwd, custom_fwd
from torch.autograd.function import Function, once_differentiable

from dfa3D import ext_loader

ext_module = ext_loader.load_ext(
    '_ext', ['wms_deform_attn_backward', 'wms_deform_attn_forward', 'ms_depth_score_sample_forward',
             'ms_depth_score_sample_backward'])


class WeightedMultiScaleDeformableAttnFunction(Function):
    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
                depth_score, im2col_step):
        ctx.im2col_step = im2col_step
        output = ext_module.wms_deform_attn_forward(
            value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
            depth_score, im2col_step=ctx.im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                              attention_weights, depth_score)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, lsi, loc, attn, ds = ctx.saved_tensors
        grad_value = torch.zeros_like(value)
        grad_sampling_loc = torch.zeros_like(loc)
        grad_attn_weight = torch.zeros_like(attn)
        grad_depth_score = torch.zeros_like(ds)
        ext_module.wms_deform_attn_backward(
            value, shapes, lsi, loc, attn, ds, grad_output.contiguous(), grad_value, grad_sampling_loc,
            grad_attn_weight, grad_depth_score, im2col_step=ctx.im2col_step)
        return grad_value, None, None, grad_sampling_loc, grad_attn_weight, grad_depth_score, None


class MultiScaleDepthScoreSampleFunction(Function):
    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, im2col_step):
        ctx.im2col_step = im2col_step
        output = ext_module.ms_depth_score_sample_forward(
            value, value_spatial_shapes, value_level_start_index, sampling_locations, im2col_step=ctx.im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, lsi, loc = ctx.saved_tensors
        grad_value = torch.zeros_like(value)
        grad_sampling_loc = torch.zeros_like(loc)
        ext_module.ms_depth_score_sample_backward(
            value, shapes, lsi, loc, grad_output.contiguous(), grad_value, grad_sampling_loc,
            im2col_step=ctx.im2col_step)
        return grad_value, None, None, grad_sampling_loc, None


class MultiScale3DDeformableAttnFunction(Function):
    """One-stage DFA3D.  Forward returns ``(output, depth_score)`` like F3D:277-302; runs the fused kernel
    (depth scores are still returned because the reference's callers read them, DCA:492)."""

    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float32)   # F3D:276 (the _fp32 variant casts under autocast)
    def forward(ctx, value, value_dpt_dist, value_spatial_shapes, value_level_start_index, sampling_locations,
                attention_weights, im2col_step):
        ctx.im2col_step = im2col_step
        output, depth_score = ext_module.dfa3d_fused_forward(
            value, value_dpt_dist, value_spatial_shapes, value_level_start_index, sampling_locations,
            attention_weights)
        ctx.save_for_backward(value, value_dpt_dist, value_spatial_shapes, value_level_start_index,
                              sampling_locations, attention_weights)
        # the reference raises NotImplementedError in backward when a gradient reaches depth_score (F3D:310-315, after a
        # host sync on its sum); here the output is declared non-differentiable, so differentiating through it fails in
        # autograd itself instead of silently dropping that gradient -- and without the sync
        ctx.mark_non_differentiable(depth_score)
        return output, depth_score

    @staticmethod
    @once_differentiable
    @custom_bwd(device_type='cuda')
    def backward(ctx, grad_output, grad_depth_score_):
        value, dist, shapes, lsi, loc, attn = ctx.saved_tensors
        grad_value = torch.zeros_like(value)
        grad_dist = torch.zeros_like(dist)
        grad_loc = torch.zeros_like(loc)
        grad_attn = torch.zeros_like(attn)
        ext_module.dfa3d_fused_backward(value, dist, shapes, lsi, loc, attn, grad_output.contiguous(), grad_value,
                                        grad_dist, grad_loc, grad_attn)
        return grad_value, grad_dist, None, None, grad_loc, grad_attn, None


MultiScale3DDeformableAttnFunction_fp32 = MultiScale3DDeformableAttnFunction
