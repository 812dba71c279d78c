from .multi_scale_3D_deform_attn import (MultiScale3DDeformableAttnFunction, MultiScaleDepthScoreSampleFunction,
                                         WeightedMultiScaleDeformableAttnFunction)

__all__ = ['MultiScaleDepthScoreSampleFunction', 'WeightedMultiScaleDeformableAttnFunction',
           'MultiScale3DDeformableAttnFunction']
