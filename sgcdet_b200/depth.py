"""The depth-distribution producer in front of the view transform, host side.

Mirrors the pieces of ``DepthNet_Fusion`` (mmdet3d_plugin/models/im2voxel/depth_utils/depth_est_fusion.py) that are
arithmetic of the hot path's neighbourhood -- neighbour selection, the homographies, the plane-sweep cost volume, the final
softmax -- and the depth pyramid of ``SGCDet.build_volume`` (detectors/SGCDet.py:83-85).  The 2-D networks between the cost
volume and the softmax (ResNetFPN, the three SimpleUnet2D) are library convolutions and out of scope; a caller runs them
with torch and hands their outputs over:

    corr  = plane_sweep_correlation(f_mvs, img_meta, stride, neighbor_img_num, depth_values)   # [V,D,H,W], fused kernel
    ...   = conv stacks on corr / mono features -> depth logits [V,D,H,W]                       # caller's torch modules
    prob, dists = depth_pyramid(logits, img_meta)            # softmax + x1, x1/2, x1/4 levels, channel-last, one kernel
    voxel_head(mlvl_feats, img_meta, dists)                  # AdaptiveSparseHead takes the DepthCL levels as they are
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import functional as SF


def get_closest_frame_ids(num_cams: int, num_select: int) -> torch.Tensor:
    """depth_est_fusion.py:52-63: the ``num_select`` temporally closest frames of every frame ([N,K] int64); at the two ends
    of the sequence the window is shifted inwards."""
    assert num_select % 2 == 0
    half = num_select // 2
    main = torch.arange(num_cams).unsqueeze(1)
    offs = torch.cat([torch.arange(-half, 0), torch.arange(1, half + 1)]).unsqueeze(0)
    ids = main + offs
    ids[0:half, :] = ids[0:half, :] + half + 1
    ids[num_cams - half:num_cams, :] = ids[num_cams - half:num_cams, :] - half - 1
    return ids


def depth_bin_centers(dbound: Sequence[float]) -> np.ndarray:
    """depth_est_fusion.py:176: centres of the depth intervals."""
    return np.arange(dbound[0], dbound[1], dbound[2], dtype=np.float32) + dbound[2] / 2


def relative_projections(img_meta: dict, stride: int, neighbor_ids: torch.Tensor) -> torch.Tensor:
    """[V,K,12] fp32 on the CPU: rows of (src_proj @ inverse(ref_proj))[:3,:3] followed by [:3,3], src = the neighbour
    frame, with the same torch ops as depth_est_fusion.py:198-207 (intrinsic rescaled to the feature map), :66-83
    (collect_proj) and :95-97 (homo_warping)."""
    w2c = torch.tensor(np.array(img_meta['lidar2img']['extrinsic']))
    intr = torch.tensor(np.array(img_meta['lidar2img']['intrinsic']))
    ratio = img_meta['ori_shape'][0] / (img_meta['img_shape'][0] / stride)
    intr = intr.clone()
    if intr.dim() == 2:
        intr[:2] /= ratio
        intr = intr.unsqueeze(0).repeat(w2c.shape[0], 1, 1)
    else:
        intr[:, :2] /= ratio
    proj = torch.matmul(intr, w2c)                                        # [V,4,4] world -> feature-map pixels
    V, K = neighbor_ids.shape
    nei = proj[neighbor_ids.reshape(-1)].view(V, K, 4, 4)
    rel = torch.matmul(nei, torch.inverse(proj).unsqueeze(1))            # src_proj @ inverse(ref_proj)
    return torch.cat([rel[:, :, :3, :3].reshape(V, K, 9), rel[:, :, :3, 3]], dim=-1).float().contiguous()


def plane_sweep_correlation(f_mvs: torch.Tensor, img_meta: dict, stride: int, neighbor_img_num: int,
                            depth_values) -> torch.Tensor:
    """``correlation`` of DepthNet_Fusion.forward (depth_est_fusion.py:209-232) for one scene: f_mvs [V,C,H,W] (CUDA, fp32) ->
    [V,D,H,W].  Differentiable w.r.t. ``f_mvs`` (the homographies carry no gradient, as in the reference)."""
    if not f_mvs.is_cuda:
        raise RuntimeError('sgcdet_b200 has no CPU implementation: inputs must be CUDA tensors')
    V = f_mvs.shape[0]
    k = min(neighbor_img_num, V - 1)
    nbr = get_closest_frame_ids(V, k)
    rt = relative_projections(img_meta, stride, nbr).to(f_mvs.device)
    depth = torch.as_tensor(np.asarray(depth_values, dtype=np.float32)).to(f_mvs.device)
    with torch.cuda.device(f_mvs.device):
        return SF.PlaneSweep.apply(f_mvs.float(), nbr.to(f_mvs.device, torch.int32).contiguous(), rt, depth)


def pyramid_crops(img_meta: dict, num_levels: int = 3) -> List[Tuple[int, int]]:
    """(h, w) the view transform crops level l to: img_shape // (4 * 2^l) (AdaptiveSparseHead.py:47-60), finest first."""
    H, W = img_meta['img_shape'][0], img_meta['img_shape'][1]
    return [(H // (4 * 2 ** l), W // (4 * 2 ** l)) for l in range(num_levels)]


def depth_pyramid(logits: torch.Tensor, img_meta: dict):
    """logits [V,D,H,W] -> (prob [1,V,D,H,W] as DepthNet_Fusion returns it, [DepthCL finest, half, quarter]) -- the list is
    what ``AdaptiveSparseHead.forward`` takes as ``mlvl_dpt_dists``."""
    if not logits.is_cuda:
        raise RuntimeError('sgcdet_b200 has no CPU implementation: inputs must be CUDA tensors')
    crops = tuple(pyramid_crops(img_meta))
    with torch.cuda.device(logits.device):
        prob, c0, c1, c2 = SF.DepthPyramid.apply(logits.float(), crops)
    return prob.unsqueeze(0), [SF.DepthCL(t, h, w) for t, (h, w) in zip((c0, c1, c2), crops)]
