"""tcgen05 feature projection (csrc/sgc_project_tc.cu) vs an fp64 reference of the same GEMM.
Tolerance: rtol 1e-3 / atol 1e-4 relative to the output scale (the bf16 hi/lo split gives ~1e-5)."""
import pytest
import torch

from sgcdet_b200 import functional as SF
from sgcdet_b200._lib import call, ptr, stream

pytestmark = pytest.mark.gpu


def run_tc(feat, h, w, wcat):
    V, C, H0, W0 = feat.shape
    N = wcat.shape[0]
    S = h * w
    assert w == W0
    wpack = torch.empty(2 * N * C, device=feat.device, dtype=torch.bfloat16)
    call('sgc_pack_weight_tc', ptr(wcat.contiguous()), N, C, ptr(wpack), stream())
    vg = torch.full((V, S, N), float('nan'), device=feat.device)
    call('sgc_project_tc_fwd', ptr(feat), C * H0 * W0, H0 * W0, V, C, S, ptr(wpack), N, ptr(vg), stream())
    return vg


@pytest.mark.parametrize('V,C,H0,W0,h,N', [
    (2, 256, 60, 80, 59, 384),    # SGCDet_ScanNet finest level (row crop, 37 tiles per view, tail tile of 112)
    (3, 128, 15, 20, 15, 256),    # "-L" coarsest level, single MMA part
    (1, 256, 5, 20, 5, 384),      # S = 100 < one tile
    (5, 256, 30, 40, 29, 384),    # many tiles, persistent loop wraps the pipelines several times
    (160, 128, 60, 80, 59, 256),  # more tiles than SMs
])
def test_project_tc_matches_fp64(cuda_lib, V, C, H0, W0, h, N):
    g = torch.Generator().manual_seed(V * 1000 + C + h)
    feat = torch.randn(V, C, H0, W0, generator=g).cuda()
    wcat = (torch.randn(N, C, generator=g) / C ** 0.5).cuda()
    vg = run_tc(feat, h, W0, wcat)
    torch.cuda.synchronize()
    ref = torch.einsum('vcs,nc->vsn', feat[:, :, :h].reshape(V, C, -1).double(), wcat.double())
    assert torch.isfinite(vg).all()
    scale = ref.abs().max().item()
    err = (vg.double() - ref).abs().max().item() / scale
    assert err < 1e-4, err
    torch.testing.assert_close(vg.double() / scale, ref / scale, rtol=1e-3, atol=1e-4)


def test_project_tc_matches_library_path(cuda_lib):
    """Same result (to fp32 round-off) as the sgc_split_bf16x3 + library bf16 GEMM path used before."""
    g = torch.Generator().manual_seed(5)
    feat = torch.randn(4, 256, 30, 40, generator=g).cuda()
    wcat = (torch.randn(384, 256, generator=g) / 16).cuda()
    a = run_tc(feat, 29, 40, wcat)
    import os
    os.environ['SGC_TC_PROJECT'] = '0'
    try:
        b = SF.ProjectFeatures.apply(feat, 29, 40, wcat)
    finally:
        os.environ.pop('SGC_TC_PROJECT')
    torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize('V,C,H0,W0,h,N', [(2, 256, 60, 80, 59, 384), (3, 128, 15, 20, 15, 256), (1, 256, 5, 20, 5, 384),
                                           (7, 256, 30, 40, 29, 384), (40, 128, 30, 40, 29, 256)])
def test_project_tc_backward_kernels_match_fp64(cuda_lib, V, C, H0, W0, h, N):
    """sgc_project_tc_bwd_data / sgc_project_tc_wgrad vs fp64 einsums of the same fp32 operands."""
    from sgcdet_b200 import _lib
    g = torch.Generator().manual_seed(V * 77 + C + h)
    S = h * W0
    feat = torch.randn(V, C, H0, W0, generator=g).cuda()
    wcat = (torch.randn(N, C, generator=g) / C ** 0.5).cuda()
    gvg = torch.randn(V, S, N, generator=g).cuda()
    wpack_t = torch.empty(2 * N * C, device='cuda', dtype=torch.bfloat16)
    call('sgc_pack_weight_tc', ptr(wcat.t().contiguous()), C, N, ptr(wpack_t), stream())
    gfeat = torch.full((V, C, H0, W0), float('nan'), device='cuda')
    gfeat[:, :, h:] = 0
    call('sgc_project_tc_bwd_data', ptr(gvg), V, S, N, ptr(wpack_t), C, ptr(gfeat), H0 * W0, stream())
    gw = torch.full((N, C), float('nan'), device='cuda')
    scratch = torch.empty(_lib.load().sgc_project_tc_wgrad_scratch_floats(N, C), device='cuda')
    call('sgc_project_tc_wgrad', ptr(gvg), ptr(feat), H0 * W0, V, S, N, C, ptr(gw), ptr(scratch), stream())
    torch.cuda.synchronize()
    ref_gf = torch.einsum('vsn,nc->vcs', gvg.double(), wcat.double())
    got_gf = gfeat[:, :, :h].reshape(V, C, S).double()
    assert torch.isfinite(gfeat).all() and torch.isfinite(gw).all()
    assert (gfeat[:, :, h:] == 0).all()
    sc1 = ref_gf.abs().max().item()
    torch.testing.assert_close(got_gf / sc1, ref_gf / sc1, rtol=1e-3, atol=1e-4)
    ref_gw = torch.einsum('vsn,vcs->nc', gvg.double(), feat[:, :, :h].reshape(V, C, S).double())
    sc2 = ref_gw.abs().max().item()
    torch.testing.assert_close(gw.double() / sc2, ref_gw / sc2, rtol=1e-3, atol=1e-4)
