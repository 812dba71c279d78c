"""Operator-level parity (boundary B2): libsgcdet_b200 through the dfa3D drop-in vs the CPU oracle.

Tolerance: rtol 1e-3 / atol 1e-4 in fp32 (BASELINE.json north_star); the oracle is evaluated in fp64 on the
same fp32 inputs so the comparison measures the kernel's error only."""
import pytest
import torch

import sgcdet_b200
from oracle import dfa3d_ref

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-3, 1e-4


def _rand_case(B, Q, M, Cm, D, shapes, P, seed, spread=1.3):
    g = torch.Generator().manual_seed(seed)
    shapes3d = torch.tensor([[h, w, D] for h, w in shapes], dtype=torch.long)
    sizes = shapes3d[:, 0] * shapes3d[:, 1]
    lsi = torch.cat([sizes.new_zeros(1), sizes.cumsum(0)[:-1]])
    S = int(sizes.sum())
    L = len(shapes)
    value = torch.randn(B, S, M, Cm, generator=g)
    dist = torch.randn(B, S, M, D, generator=g).softmax(-1)
    # locations spill outside [0,1] so every border / out-of-range branch is exercised
    loc = (torch.rand(B, Q, M, L, P, 3, generator=g) - 0.5) * spread + 0.5
    attn = torch.rand(B, Q, M, L, P, generator=g)
    gout = torch.randn(B, Q, M * Cm, generator=g)
    return value, dist, shapes3d, lsi, loc, attn, gout


CASES = [
    # B, Q, M, Cm, D, shapes, P
    (2, 37, 8, 32, 12, [(7, 9)], 4),           # SGCDet stage-2 shape family
    (3, 50, 1, 256, 12, [(5, 8)], 1),          # SGCDet stage-1 (Grid_Sample_3D_Feature) shape family
    (2, 33, 8, 16, 12, [(6, 8)], 4),           # "-L" configs (Cm = 16)
    (2, 41, 8, 32, 28, [(12, 20), (6, 10), (3, 5), (2, 3)], 8),  # unittest_DFA3D.py family: 4 levels, P = 8
    (1, 5, 3, 5, 4, [(3, 2)], 2),              # odd sizes, Cm not a multiple of 4
]


def _ops():
    sgcdet_b200.install_dropin()
    from dfa3D.ops import (MultiScale3DDeformableAttnFunction, MultiScaleDepthScoreSampleFunction,
                           WeightedMultiScaleDeformableAttnFunction)
    return MultiScaleDepthScoreSampleFunction, WeightedMultiScaleDeformableAttnFunction, MultiScale3DDeformableAttnFunction


@pytest.mark.parametrize('case', CASES)
def test_two_stage_forward_backward(cuda_lib, case):
    DS, WMS, _ = _ops()
    value, dist, shapes3d, lsi, loc, attn, gout = _rand_case(*case, seed=11)
    dev = 'cuda'
    v, d, l, a = (t.to(dev).requires_grad_(True) for t in (value, dist, loc, attn))
    s3, ls = shapes3d.to(dev), lsi.to(dev)
    ds = DS.apply(d, s3, ls, l, 64)
    out = WMS.apply(v, s3[:, :2].contiguous(), ls, l[..., :2].contiguous(), a, ds, 64)
    # oracle in fp64
    v64, d64, l64, a64 = (t.double().requires_grad_(True) for t in (value, dist, loc, attn))
    ds_ref = dfa3d_ref.depth_score_sample_forward(d64, shapes3d, lsi, l64)
    out_ref = dfa3d_ref.wms_deform_attn_forward(v64, shapes3d[:, :2], lsi, l64[..., :2], a64, ds_ref)
    torch.testing.assert_close(ds.cpu().double(), ds_ref.detach(), rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(out.cpu().double(), out_ref.detach(), rtol=RTOL, atol=ATOL)
    out.backward(gout.to(dev))
    out_ref.backward(gout.double())
    for name, got, ref in (('value', v, v64), ('dist', d, d64), ('loc', l, l64), ('attn', a, a64)):
        torch.testing.assert_close(got.grad.cpu().double(), ref.grad, rtol=RTOL, atol=ATOL, msg=lambda m: f'{name}: {m}')


@pytest.mark.parametrize('case', CASES)
def test_one_stage_equals_two_stage_and_oracle(cuda_lib, case):
    """unittest_DFA3D.py:11-28 computes both and never compares them; here they must agree."""
    DS, WMS, F3 = _ops()
    value, dist, shapes3d, lsi, loc, attn, gout = _rand_case(*case, seed=5)
    dev = 'cuda'
    v, d, l, a = (t.to(dev).requires_grad_(True) for t in (value, dist, loc, attn))
    s3, ls = shapes3d.to(dev), lsi.to(dev)
    out1, ds1 = F3.apply(v, d, s3, ls, l, a, 64)
    ds2 = DS.apply(d.detach(), s3, ls, l.detach(), 64)
    out2 = WMS.apply(v.detach(), s3[:, :2].contiguous(), ls, l.detach()[..., :2].contiguous(), a.detach(), ds2, 64)
    assert torch.equal(ds1, ds2)
    torch.testing.assert_close(out1, out2, rtol=1e-6, atol=1e-6)
    out1.backward(gout.to(dev))
    g_value, g_dist, g_loc, g_attn = dfa3d_ref.dfa3d_backward(value.double(), dist.double(), shapes3d, lsi,
                                                              loc.double(), attn.double(), gout.double())
    for name, got, ref in (('value', v, g_value), ('dist', d, g_dist), ('loc', l, g_loc), ('attn', a, g_attn)):
        torch.testing.assert_close(got.grad.cpu().double(), ref, rtol=RTOL, atol=ATOL, msg=lambda m: f'{name}: {m}')


def test_contract_errors(cuda_lib):
    """Error behaviour of the reference launchers (WMSL:220-253): CPU / non-contiguous tensors raise."""
    sgcdet_b200.install_dropin()
    from dfa3D import ext_loader
    ext = ext_loader.load_ext('_ext', ['wms_deform_attn_backward', 'wms_deform_attn_forward',
                                       'ms_depth_score_sample_forward', 'ms_depth_score_sample_backward'])
    value, dist, shapes3d, lsi, loc, attn, _ = _rand_case(2, 4, 2, 8, 6, [(4, 4)], 2, seed=1)
    with pytest.raises(RuntimeError):
        ext.ms_depth_score_sample_forward(dist, shapes3d, lsi, loc, im2col_step=64)  # CPU tensors
    dc, lc = dist.cuda(), loc.cuda()
    with pytest.raises(RuntimeError):
        ext.ms_depth_score_sample_forward(dc.transpose(0, 1), shapes3d.cuda(), lsi.cuda(), lc, im2col_step=64)
    with pytest.raises(RuntimeError):  # batch 3 % min(3,2) != 0
        v3 = torch.randn(3, 16, 2, 6).softmax(-1).cuda()
        ext.ms_depth_score_sample_forward(v3, shapes3d.cuda(), lsi.cuda(), torch.rand(3, 4, 2, 1, 2, 3).cuda(), im2col_step=2)


def test_empty_queries(cuda_lib):
    DS, WMS, F3 = _ops()
    value, dist, shapes3d, lsi, loc, attn, _ = _rand_case(2, 0, 2, 8, 6, [(4, 4)], 2, seed=1)
    out, ds = F3.apply(value.cuda(), dist.cuda(), shapes3d.cuda(), lsi.cuda(), loc.cuda(), attn.cuda(), 64)
    assert out.shape == (2, 0, 16) and ds.shape == (2, 0, 2, 1, 2, 4)


def test_corner_order_known_answer(cuda_lib):
    """KAT for the depth_score corner order [TL, TR, BR, BL] (DSK:89-92) and loc order (x=w, y=h, z=d)."""
    DS, _, _ = _ops()
    H, W, D = 2, 2, 1
    dist = torch.tensor([1.0, 2.0, 3.0, 4.0]).view(1, 4, 1, 1)  # pixel (h,w) -> 1 + 2h + w
    shapes3d = torch.tensor([[H, W, D]])
    lsi = torch.zeros(1, dtype=torch.long)
    loc = torch.tensor([0.5, 0.5, 0.5]).view(1, 1, 1, 1, 1, 3)  # centre: h_im = w_im = 0.5, d_im = 0
    ds = DS.apply(dist.cuda(), shapes3d.cuda(), lsi.cuda(), loc.cuda(), 64).cpu().view(4)
    assert ds.tolist() == [1.0, 2.0, 4.0, 3.0]
