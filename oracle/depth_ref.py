"""CPU restatement of the depth-distribution producer's arithmetic (test infrastructure only).

Reference files restated here (same torch functions the reference calls, on the CPU):
  * ``mmdet3d_plugin/models/im2voxel/depth_utils/depth_est_fusion.py:52-63``   get_closest_frame_ids
  * ``.../depth_est_fusion.py:66-83``                                         collect_proj
  * ``.../depth_est_fusion.py:85-126``                                        homo_warping (plane-sweep homography + grid_sample)
  * ``.../depth_est_fusion.py:198-232``                                       intrinsic rescale, neighbour loop, correlation
  * ``.../depth_est_fusion.py:241``                                           softmax over the depth bins
  * ``mmdet3d_plugin/models/detectors/SGCDet.py:83-85``                       nearest x1/2, x1/4 depth pyramid
  * ``mmdet3d_plugin/models/im2voxel/AdaptiveSparseHead.py:47-60`` + ``transformer_utils/transformer.py:151-170``
                                                                              the [:h,:w] crop and the channel-last flatten

Pinning: ``tests/golden/make_golden_depth.py`` executes the reference's OWN ``get_closest_frame_ids`` / ``collect_proj`` /
``homo_warping`` (function definitions taken from /root/reference at run time, nothing copied) on seeded inputs and
freezes neighbour ids, warped features and the correlation as ``tests/golden/depth_producer.pt``;
``tests/test_oracle_cpu.py::test_depth_oracle_matches_reference_golden`` checks this file against it.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import torch
import torch.nn.functional as F


def closest_frame_ids(num_cams: int, num_select: int) -> torch.Tensor:
    """depth_est_fusion.py:52-63."""
    assert num_select % 2 == 0
    half = num_select // 2
    main = torch.arange(num_cams).unsqueeze(1)
    offsets = torch.cat([torch.arange(-half, 0).unsqueeze(0), torch.arange(1, half + 1).unsqueeze(0)], dim=1)
    closest = main + offsets
    closest[0:half, :] = closest[0:half, :] + half + 1
    closest[num_cams - half:num_cams, :] = closest[num_cams - half:num_cams, :] - half - 1
    return closest


def feature_intrinsic(intrinsic: torch.Tensor, img_shape, ori_shape, stride: int) -> torch.Tensor:
    """depth_est_fusion.py:198-207: the intrinsic rescaled to the feature map's resolution."""
    ratio = ori_shape[0] / (img_shape[0] / stride)
    k = intrinsic.clone()
    if k.dim() == 2:
        k[:2] /= ratio
    else:
        k[:, :2] /= ratio
    return k


def collect_proj(w2c: torch.Tensor, intr: torch.Tensor, neighbor_ids: torch.Tensor):
    """depth_est_fusion.py:66-83 -> (proj [V,4,4], [K x (V,4,4)])."""
    if intr.dim() == 2:
        intr = intr.unsqueeze(0).repeat(w2c.shape[0], 1, 1)
    proj = torch.matmul(intr, w2c)
    V, K = neighbor_ids.shape
    nei = proj[neighbor_ids.reshape(-1)].view(V, K, 4, 4)
    return proj, list(torch.unbind(nei, dim=1))


def homo_warp(src_fea: torch.Tensor, src_proj: torch.Tensor, ref_proj: torch.Tensor, depth_values: torch.Tensor):
    """depth_est_fusion.py:85-126 -> warped [B,C,D,H,W]."""
    B, C, H, W = src_fea.shape
    D = depth_values.shape[1]
    with torch.no_grad():
        proj = torch.matmul(src_proj, torch.inverse(ref_proj))
        rot, trans = proj[:, :3, :3], proj[:, :3, 3:4]
        y, x = torch.meshgrid([torch.arange(0, H, dtype=torch.float32, device=src_fea.device),
                               torch.arange(0, W, dtype=torch.float32, device=src_fea.device)], indexing='ij')
        y, x = y.contiguous().view(H * W), x.contiguous().view(H * W)
        xyz = torch.stack((x, y, torch.ones_like(x))).unsqueeze(0).repeat(B, 1, 1)
        rot_xyz = torch.matmul(rot, xyz)
        rot_depth_xyz = rot_xyz.unsqueeze(2).repeat(1, 1, D, 1) * depth_values.view(B, 1, D, 1)
        proj_xyz = rot_depth_xyz + trans.view(B, 3, 1, 1)
        proj_xy = proj_xyz[:, :2] / proj_xyz[:, 2:3]
        gx = proj_xy[:, 0] / ((W - 1) / 2) - 1
        gy = proj_xy[:, 1] / ((H - 1) / 2) - 1
        grid = torch.stack((gx, gy), dim=3)
    warped = F.grid_sample(src_fea, grid.view(B, D * H, W, 2), mode='bilinear', padding_mode='zeros', align_corners=False)
    return warped.view(B, C, D, H, W)


def plane_sweep_correlation(f_mvs: torch.Tensor, w2c: torch.Tensor, intr_feat: torch.Tensor, depth_values: torch.Tensor,
                            neighbor_img_num: int) -> torch.Tensor:
    """depth_est_fusion.py:209-232: f_mvs [V,C,H,W] -> correlation [V,D,H,W]."""
    V, C, H, W = f_mvs.shape
    k = min(neighbor_img_num, V - 1)
    ids = closest_frame_ids(V, k).to(f_mvs.device)
    nei_feats = torch.unbind(f_mvs[ids.view(-1)].view(V, k, C, H, W), dim=1)
    ref_proj, nei_projs = collect_proj(w2c, intr_feat, ids)
    dv = depth_values.view(1, -1).repeat(V, 1)
    corr = torch.zeros(V, dv.shape[1], H, W, dtype=f_mvs.dtype, device=f_mvs.device)
    for nf, npj in zip(nei_feats, nei_projs):
        warped = homo_warp(nf, npj, ref_proj, dv)
        corr = corr + (warped * f_mvs.unsqueeze(2)).sum(dim=1) / math.sqrt(float(C))
    return corr / k


def depth_pyramid(logits: torch.Tensor, crops: Sequence[Tuple[int, int]]):
    """depth_est_fusion.py:241 + SGCDet.py:83-85 + the crop / channel-last flatten of the view transform:
    logits [V,D,H,W] -> (prob [V,D,H,W], [level l: [V, h_l*w_l, D]])."""
    prob = F.softmax(logits, dim=1)
    p5 = prob.unsqueeze(0)                                            # [B=1, N, D, H, W]
    lv = [p5, F.interpolate(p5, scale_factor=(1, 0.5, 0.5), mode='nearest'),
          F.interpolate(p5, scale_factor=(1, 0.25, 0.25), mode='nearest')]
    out: List[torch.Tensor] = []
    for t, (h, w) in zip(lv, crops):
        out.append(t[0, :, :, :h, :w].permute(0, 2, 3, 1).reshape(t.shape[1], h * w, -1).contiguous())
    return prob, out
