"""Deterministic synthetic scenes and the shape table of the shipped SGCDet configs.

Shared by tests and bench.py.  Everything is generated on the CPU with a seeded
``torch.Generator`` (default seed 1234 = the reference's default, ``main.py:20``) so the CPU
oracle and the GPU path see identical bits; callers copy to the device.

Shapes follow ``configs/SGCDet_{ScanNet,ARKit,large_ScanNet200,large_ARKit}.py:1-15`` and the
``img_meta`` wire format of ``mmdet3d_plugin/datasets/scannet_multiview_dataset.py:29-38``
(``lidar2img = {intrinsic 4x4, extrinsic [V x 4x4 world->cam], origin [3]}``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np
import torch


@dataclass(frozen=True)
class PathConfig:
    """The hot-path hyper-parameters one ``SGCDet_*.py`` config fixes."""
    name: str
    embed_dims: int
    n_voxels_list: Tuple[Tuple[int, int, int], ...]
    voxel_size_list: Tuple[Tuple[float, float, float], ...]
    topk_list: Tuple[int, ...]
    img_shape: Tuple[int, int]
    ori_shape: Tuple[int, int]
    intrinsic: Tuple[float, float, float, float]  # fx, fy, cx, cy
    dbound: Tuple[float, float, float] = (0.2, 5.0, 0.4)
    num_heads: int = 8
    num_points: int = 4
    depth_bins: int = 12
    feat_hw: Tuple[Tuple[int, int], ...] = ((60, 80), (30, 40), (15, 20))  # stride 4, 8, 16 maps

    @property
    def num_levels(self) -> int:
        return len(self.n_voxels_list)

    def level_hw(self, i: int) -> Tuple[int, int]:
        """Cropped feature-map size used by head level i (AdaptiveSparseHead.py:51-59)."""
        stride = 4 * 2 ** (self.num_levels - 1 - i)
        return self.img_shape[0] // stride, self.img_shape[1] // stride

    def level_queries(self, i: int) -> int:
        n = int(np.prod(self.n_voxels_list[i]))
        return n if i == 0 else min(n, self.topk_list[i - 1])


_SCANNET_K = (1169.6, 1167.1, 646.3, 489.9)
_ARKIT_K = (212.0, 212.0, 128.0, 96.0)

CONFIGS: Dict[str, PathConfig] = {
    'SGCDet_ScanNet': PathConfig(
        'SGCDet_ScanNet', 256, ((10, 10, 4), (20, 20, 8), (40, 40, 16)),
        ((.64, .64, .8), (.32, .32, .4), (.16, .16, .2)), (800, 6400),
        (239, 320), (968, 1296), _SCANNET_K),
    'SGCDet_ARKit': PathConfig(
        'SGCDet_ARKit', 256, ((10, 10, 4), (20, 20, 8), (40, 40, 16)),
        ((.64, .64, .8), (.32, .32, .4), (.16, .16, .2)), (800, 6400),
        (240, 320), (192, 256), _ARKIT_K),
    'SGCDet_large_ScanNet200': PathConfig(
        'SGCDet_large_ScanNet200', 128, ((20, 20, 8), (40, 40, 16), (80, 80, 32)),
        ((.32, .32, .4), (.16, .16, .2), (.08, .08, .1)), (6400, 51200),
        (239, 320), (968, 1296), _SCANNET_K),
    'SGCDet_large_ARKit': PathConfig(
        'SGCDet_large_ARKit', 128, ((20, 20, 8), (40, 40, 16), (80, 80, 32)),
        ((.32, .32, .4), (.16, .16, .2), (.08, .08, .1)), (6400, 51200),
        (240, 320), (192, 256), _ARKIT_K),
    # reduced shape for fast CPU/GPU parity tests (same structure, 3 levels)
    'tiny': PathConfig(
        'tiny', 128, ((4, 4, 2), (8, 8, 4), (16, 16, 8)),
        ((1.6, 1.6, 1.6), (.8, .8, .8), (.4, .4, .4)), (64, 512),
        (95, 128), (968, 1296), _SCANNET_K, feat_hw=((24, 32), (12, 16), (6, 8))),
    # the tiny grids with the channel count of the C=256 configs (32-wide heads: the tensor-core per-head products)
    'tiny256': PathConfig(
        'tiny256', 256, ((4, 4, 2), (8, 8, 4), (16, 16, 8)),
        ((1.6, 1.6, 1.6), (.8, .8, .8), (.4, .4, .4)), (64, 512),
        (95, 128), (968, 1296), _SCANNET_K, feat_hw=((24, 32), (12, 16), (6, 8))),
}


def _extrinsics(gen: torch.Generator, num_views: int) -> List[np.ndarray]:
    """world->cam 4x4 (cam x=right, y=down, z=forward; world z up)."""
    u = torch.rand(num_views, 5, generator=gen, dtype=torch.float64)
    out = []
    for i in range(num_views):
        px = -2.5 + 5.0 * u[i, 0].item()
        py = -2.5 + 5.0 * u[i, 1].item()
        pz = 1.0 + 0.8 * u[i, 2].item()
        yaw = 2 * math.pi * u[i, 3].item()
        pitch = math.radians(-30.0 + 40.0 * u[i, 4].item())
        f = np.array([math.cos(pitch) * math.cos(yaw), math.cos(pitch) * math.sin(yaw), math.sin(pitch)])
        r = np.array([math.sin(yaw), -math.cos(yaw), 0.0])
        d = np.cross(f, r)
        R = np.stack([r, d, f])
        E = np.eye(4)
        E[:3, :3] = R
        E[:3, 3] = -R @ np.array([px, py, pz])
        out.append(E.astype(np.float32))
    return out


def make_img_meta(cfg: PathConfig, num_views: int, gen: torch.Generator, shift_origin: bool = False) -> dict:
    fx, fy, cx, cy = cfg.intrinsic
    K = np.eye(4, dtype=np.float32)
    K[0, 0], K[1, 1], K[0, 2], K[1, 2] = fx, fy, cx, cy
    origin = np.array([0.0, 0.0, 0.5], dtype=np.float32)
    ext = _extrinsics(gen, num_views)
    if shift_origin:  # configs/SGCDet_ScanNet.py:147, pipelines/multi_view.py:8-15
        s = torch.randn(3, generator=gen, dtype=torch.float64).numpy() * np.array([0.7, 0.7, 0.0])
        origin = (origin + s).astype(np.float32)
    return dict(
        img_shape=(cfg.img_shape[0], cfg.img_shape[1], 3),
        ori_shape=(cfg.ori_shape[0], cfg.ori_shape[1], 3),
        lidar2img=dict(intrinsic=K, extrinsic=ext, origin=origin),
    )


@dataclass
class Scene:
    cfg: PathConfig
    num_views: int
    img_meta: dict
    mlvl_feats: List[torch.Tensor]       # [1,V,C,h,w] fp32, finest first (as the FPN gives them)
    mlvl_dpt_dists: List[torch.Tensor]   # [1,V,D,h,w]
    grad_volume: torch.Tensor            # upstream gradient [1,C,X,Y,Z]
    geo_occ: torch.Tensor                # [1, sum N] Bernoulli(0.2) occupancy GT for occ_loss
    extra: dict = field(default_factory=dict)

    def to(self, device) -> 'Scene':
        return Scene(self.cfg, self.num_views, self.img_meta,
                     [t.to(device) for t in self.mlvl_feats],
                     [t.to(device) for t in self.mlvl_dpt_dists],
                     self.grad_volume.to(device), self.geo_occ.to(device), self.extra)


def make_scene(cfg: PathConfig | str = 'SGCDet_ScanNet', num_views: int = 40, seed: int = 1234,
               shift_origin: bool = False) -> Scene:
    """Synthetic scene of SURVEY.md section 8d."""
    if isinstance(cfg, str):
        cfg = CONFIGS[cfg]
    gen = torch.Generator(device='cpu').manual_seed(seed)
    meta = make_img_meta(cfg, num_views, gen, shift_origin)
    C, D, V = cfg.embed_dims, cfg.depth_bins, num_views
    feats = [torch.randn(1, V, C, h, w, generator=gen) for (h, w) in cfg.feat_hw]
    h0, w0 = cfg.feat_hw[0]
    dpt = torch.softmax(2.0 * torch.randn(1, V, D, h0, w0, generator=gen), dim=2)
    # SGCDet.py:83-85: nearest /2, /4 of the depth distribution
    dists = [dpt,
             torch.nn.functional.interpolate(dpt, scale_factor=(1, 0.5, 0.5), mode='nearest'),
             torch.nn.functional.interpolate(dpt, scale_factor=(1, 0.25, 0.25), mode='nearest')]
    X, Y, Z = cfg.n_voxels_list[-1]
    grad = torch.randn(1, C, X, Y, Z, generator=gen)
    n_occ = sum(int(np.prod(n)) for n in cfg.n_voxels_list[1:])
    geo = (torch.rand(1, n_occ, generator=gen) < 0.2).float()
    return Scene(cfg, V, meta, feats, dists, grad, geo)


# ---- weights -------------------------------------------------------------------------------

def level_param_shapes(C: int, M: int = 8, P: int = 4) -> Dict[str, Tuple[int, ...]]:
    """Parameter names/shapes of ONE DenseHead level, relative to
    ``base_heads.{i}.cross_transformer.encoder.layers.0.``."""
    da = 'attentions.0.deformable_attention.'
    return {
        da + 'sampling_offsets.weight': (M * P * 2, C), da + 'sampling_offsets.bias': (M * P * 2,),
        da + 'sampling_offsets_depth.weight': (M * P, C), da + 'sampling_offsets_depth.bias': (M * P,),
        da + 'attention_weights.weight': (M * P, C), da + 'attention_weights.bias': (M * P,),
        da + 'value_proj.weight': (C, C), da + 'value_proj.bias': (C,),
        'attentions.0.output_proj.weight': (C, C), 'attentions.0.output_proj.bias': (C,),
        'attentions.0.attention_pooling.in_proj_weight': (3 * C, C),
        'attentions.0.attention_pooling.in_proj_bias': (3 * C,),
        'attentions.0.attention_pooling.out_proj.weight': (C, C),
        'attentions.0.attention_pooling.out_proj.bias': (C,),
        'ffns.0.layers.0.0.weight': (2 * C, C), 'ffns.0.layers.0.0.bias': (2 * C,),
        'ffns.0.layers.1.weight': (C, 2 * C), 'ffns.0.layers.1.bias': (C,),
        'norms.0.weight': (C,), 'norms.0.bias': (C,),
        'norms.1.weight': (C,), 'norms.1.bias': (C,),
    }


def _ring_bias(M: int, P: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Reference init of the offset biases (DCA:194-208, 351-362), num_levels=1."""
    thetas = torch.arange(M, dtype=torch.float32) * (2.0 * math.pi / M)
    g = torch.stack([thetas.cos(), thetas.sin()], -1)
    g = (g / g.abs().max(-1, keepdim=True)[0]).view(M, 1, 1, 2).repeat(1, 1, P, 1)
    gd = (((thetas.cos() + thetas.sin()) / 2).view(M, 1, 1, 1)).repeat(1, 1, P, 1)
    for i in range(P):
        g[:, :, i, :] *= i + 1
        gd[:, :, i, :] *= i + 1
    return g.reshape(-1), gd.reshape(-1)


def make_state_dict(cfg: PathConfig | str, seed: int = 4321) -> Dict[str, torch.Tensor]:
    """Random, non-degenerate weights under the reference's state-dict keys (prefix ``voxel_head.``
    stripped): every Linear/MHA weight is N(0, s^2) and every bias is small and non-zero so that
    offsets span a few pixels and attention weights are not uniform (the reference's default init
    zeroes ``sampling_offsets.weight`` and ``attention_weights``, which would hide bugs)."""
    if isinstance(cfg, str):
        cfg = CONFIGS[cfg]
    C, M, P = cfg.embed_dims, cfg.num_heads, cfg.num_points
    gen = torch.Generator(device='cpu').manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    ring_uv, ring_d = _ring_bias(M, P)
    for i in range(cfg.num_levels):
        pre = f'base_heads.{i}.cross_transformer.encoder.layers.0.'
        for k, shp in level_param_shapes(C, M, P).items():
            if k.startswith('norms') and k.endswith('weight'):
                t = 1.0 + 0.1 * torch.randn(shp, generator=gen)
            elif k.endswith('bias'):
                t = 0.05 * torch.randn(shp, generator=gen)
                if 'sampling_offsets.bias' in k:
                    t = t + ring_uv
                elif 'sampling_offsets_depth.bias' in k:
                    t = t + 0.5 * ring_d
            else:
                fan_in = shp[-1]
                gain = 1.0
                if 'sampling_offsets' in k or 'attention_weights' in k:
                    gain = 8.0  # the sampled query is small (depth-weighted average)
                t = gain / math.sqrt(fan_in) * torch.randn(shp, generator=gen)
            sd[pre + k] = t
        # DenseHead buffers (DenseHead.py:32-48)
        n = torch.tensor(cfg.n_voxels_list[i])
        vs = torch.tensor(cfg.voxel_size_list[i])
        xv, yv, zv = torch.meshgrid(torch.arange(n[0]), torch.arange(n[1]), torch.arange(n[2]), indexing='ij')
        idx = torch.arange(int(n.prod()))
        sd[f'base_heads.{i}.vox_coords'] = torch.cat(
            [xv.reshape(-1, 1), yv.reshape(-1, 1), zv.reshape(-1, 1), idx.reshape(-1, 1)], dim=-1)
        pts = torch.stack([xv, yv, zv]) * vs.view(3, 1, 1, 1) + (-n / 2. * vs).view(3, 1, 1, 1)
        sd[f'base_heads.{i}.ref_3d'] = pts.view(3, -1).permute(1, 0).contiguous()
    for i in range(cfg.num_levels - 1):
        sd[f'occ_pred_heads.{i}.0.weight'] = 1.0 / math.sqrt(C) * torch.randn(1, C, generator=gen)
        sd[f'occ_pred_heads.{i}.0.bias'] = 0.05 * torch.randn(1, generator=gen)
    return sd


# ---- roofline accounting (SURVEY.md section 8d / BASELINE.md section 4) -----------------------

def algorithmic_bytes(cfg: PathConfig | str, num_views: int) -> Dict[str, float]:
    if isinstance(cfg, str):
        cfg = CONFIGS[cfg]
    C, D, V = cfg.embed_dims, cfg.depth_bins, num_views
    w_level = 4 * (6 * C * C + 128 * C + 4 * C * C + 8 * C)
    Wb = cfg.num_levels * w_level
    fwd = bwd = 0.0
    for i in range(cfg.num_levels):
        h, w = cfg.level_hw(i)
        S = h * w
        N = int(np.prod(cfg.n_voxels_list[i]))
        fwd += 4 * (V * S * (C + D) + 2 * V * S * C + 2 * N * C)
        bwd += 4 * (V * S * (2 * C + D) + 2 * V * S * C + V * S * (C + D) + 2 * N * C)
    return dict(fwd=fwd + Wb, bwd=bwd + 2 * Wb, fwd_bwd=fwd + bwd + 3 * Wb)
