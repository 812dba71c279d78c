"""The depth-distribution producer (csrc/sgc_depth.cu, sgcdet_b200/depth.py) against the CPU oracle
(oracle/depth_ref.py, pinned to the reference's own functions by tests/golden/depth_producer.pt): plane-sweep correlation
forward + feature gradient, softmax + pyramid forward + backward, the transposes, and the hand-off of the channel-last
levels to AdaptiveSparseHead."""
import os

import numpy as np
import pytest
import torch

from oracle import depth_ref
from sgcdet_b200 import depth as SD, functional as SF, plugin, synthetic as syn
from sgcdet_b200._lib import call, ptr, stream

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-3, 1e-4
DEV = 'cuda'


def _golden():
    return torch.load(os.path.join(os.path.dirname(__file__), 'golden', 'depth_producer.pt'))


def _meta_from_case(c, H, W, stride=4):
    """img_meta whose feature-map intrinsic equals the case's: ori_shape == img_shape / stride * ... -> ratio 1 at `stride`."""
    K = c['intr_feat'].numpy().copy()
    return dict(img_shape=(H * stride, W * stride, 3), ori_shape=(H, W, 3),
                lidar2img=dict(intrinsic=K, extrinsic=[e.numpy() for e in c['w2c']]))


@pytest.mark.parametrize('name', ['a', 'b', 'c'])
def test_plane_sweep_matches_reference_golden(cuda_lib, name):
    c = _golden()[name]
    V, C, H, W = c['f_mvs'].shape
    f = c['f_mvs'].to(DEV).requires_grad_(True)
    corr = SD.plane_sweep_correlation(f, _meta_from_case(c, H, W), 4, c['k'], c['depth'].numpy())
    torch.testing.assert_close(corr.detach().cpu(), c['correlation'], rtol=RTOL, atol=ATOL)
    # gradient w.r.t. the matching features (both roles: reference pixel and warped neighbour) vs autograd of the oracle
    g = torch.Generator().manual_seed(5)
    gc = torch.randn(corr.shape, generator=g)
    corr.backward(gc.to(DEV))
    fo = c['f_mvs'].clone().requires_grad_(True)
    depth_ref.plane_sweep_correlation(fo, c['w2c'], c['intr_feat'], c['depth'], c['k']).backward(gc)
    torch.testing.assert_close(f.grad.cpu(), fo.grad, rtol=RTOL, atol=ATOL)


def test_plane_sweep_config_shape_against_oracle(cuda_lib):
    """C = 128 matching features at a (reduced-view) ScanNet feature-map shape, K = 2, D = 12 (configs/SGCDet_ScanNet.py:3,91)."""
    cfg = syn.CONFIGS['SGCDet_ScanNet']
    V, C, H, W = 6, 128, 60, 80
    g = torch.Generator().manual_seed(21)
    f = torch.randn(V, C, H, W, generator=g)
    meta = syn.make_img_meta(cfg, V, g)
    # a video-like trajectory: small steps between consecutive frames so that the neighbours overlap
    base = meta['lidar2img']['extrinsic'][0]
    ext = []
    for i in range(V):
        T = np.eye(4, dtype=np.float32)
        T[0, 3], T[2, 3] = 0.05 * i, -0.03 * i
        ext.append((T @ base).astype(np.float32))
    meta['lidar2img']['extrinsic'] = ext
    depth = SD.depth_bin_centers(cfg.dbound)
    assert depth.shape == (12,)
    fg = f.to(DEV).requires_grad_(True)
    corr = SD.plane_sweep_correlation(fg, meta, 4, 2, depth)
    intr = depth_ref.feature_intrinsic(torch.tensor(meta['lidar2img']['intrinsic']), meta['img_shape'], meta['ori_shape'], 4)
    fo = f.clone().requires_grad_(True)
    ref = depth_ref.plane_sweep_correlation(fo, torch.tensor(np.stack(ext)), intr, torch.tensor(depth), 2)
    assert float((ref != 0).float().mean()) > 0.5          # the sweep actually samples inside the neighbour images
    torch.testing.assert_close(corr.detach().cpu(), ref.detach(), rtol=RTOL, atol=ATOL)
    gc = torch.randn(ref.shape, generator=g)
    corr.backward(gc.to(DEV))
    ref.backward(gc)
    torch.testing.assert_close(fg.grad.cpu(), fo.grad, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize('V,D,H,W,crops', [(3, 12, 60, 80, ((59, 80), (29, 40), (14, 20))), (2, 5, 13, 18, ((13, 18), (6, 9), (3, 4))),
                                            (1, 32, 8, 8, ((8, 8), (4, 4), (2, 2)))])
def test_depth_pyramid_forward_backward(cuda_lib, V, D, H, W, crops):
    g = torch.Generator().manual_seed(9)
    logits = torch.randn(V, D, H, W, generator=g) * 2
    lg = logits.to(DEV).requires_grad_(True)
    prob, c0, c1, c2 = SF.DepthPyramid.apply(lg, crops)
    lo = logits.clone().requires_grad_(True)
    prob_r, lv_r = depth_ref.depth_pyramid(lo, crops)
    torch.testing.assert_close(prob.detach().cpu(), prob_r.detach(), rtol=1e-5, atol=1e-6)
    for a, b in zip((c0, c1, c2), lv_r):
        torch.testing.assert_close(a.detach().cpu(), b.detach(), rtol=1e-5, atol=1e-6)
    gs = [torch.randn(t.shape, generator=g) for t in (prob_r, *lv_r)]
    torch.autograd.backward([prob, c0, c1, c2], [t.to(DEV) for t in gs])
    torch.autograd.backward([prob_r, *lv_r], gs)
    torch.testing.assert_close(lg.grad.cpu(), lo.grad, rtol=1e-4, atol=1e-6)
    # only some outputs used (the depth loss off, or a level unused): missing gradients are zeros
    lg2 = logits.to(DEV).requires_grad_(True)
    out = SF.DepthPyramid.apply(lg2, crops)
    out[2].backward(gs[2].to(DEV))
    lo2 = logits.clone().requires_grad_(True)
    depth_ref.depth_pyramid(lo2, crops)[1][1].backward(gs[2])
    torch.testing.assert_close(lg2.grad.cpu(), lo2.grad, rtol=1e-4, atol=1e-6)


def test_transposes_roundtrip(cuda_lib):
    x = torch.randn(3, 20, 77, device=DEV)
    y = torch.empty(3, 77, 20, device=DEV)
    call('sgc_nchw_to_nhwc', ptr(x), 3, 20, 77, ptr(y), stream())
    assert torch.equal(y, x.transpose(1, 2).contiguous())
    z = torch.empty_like(x)
    call('sgc_nhwc_to_nchw', ptr(y), 3, 20, 77, ptr(z), stream())
    assert torch.equal(z, x)


def test_head_takes_channel_last_depth_levels(cuda_lib):
    """AdaptiveSparseHead fed with depth_pyramid()'s DepthCL levels == fed with the reference-layout [1,V,D,H,W] pyramid of
    the same probabilities (SGCDet.py:83-85), forward and the gradient w.r.t. the depth logits."""
    cfg = syn.CONFIGS['tiny']
    V = 6
    sc = syn.make_scene(cfg, V, shift_origin=True).to(DEV)
    head = plugin.build_voxel_head(cfg)
    head.load_state_dict(syn.make_state_dict(cfg))
    head = head.to(DEV).eval()
    D = sc.mlvl_dpt_dists[0].shape[2]
    H, W = sc.mlvl_dpt_dists[0].shape[-2:]
    logits = torch.randn(V, D, H, W, device=DEV, generator=torch.Generator(device=DEV).manual_seed(2))
    la = logits.clone().requires_grad_(True)
    prob, levels = SD.depth_pyramid(la, sc.img_meta)
    vol_a, valid_a, occ_a = head(sc.mlvl_feats, sc.img_meta, levels)
    (vol_a * sc.grad_volume).sum().backward()
    lb = logits.clone().requires_grad_(True)
    pb = torch.softmax(lb, dim=1).unsqueeze(0)
    pyr = [pb, torch.nn.functional.interpolate(pb, scale_factor=(1, 0.5, 0.5), mode='nearest'),
           torch.nn.functional.interpolate(pb, scale_factor=(1, 0.25, 0.25), mode='nearest')]
    vol_b, valid_b, occ_b = head(sc.mlvl_feats, sc.img_meta, pyr)
    (vol_b * sc.grad_volume).sum().backward()
    assert torch.equal(valid_a, valid_b)
    torch.testing.assert_close(vol_a, vol_b, rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(la.grad, lb.grad, rtol=RTOL, atol=ATOL)
