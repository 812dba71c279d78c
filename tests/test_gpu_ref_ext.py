"""Pins the oracle (and the product) against the reference's OWN kernels: the unmodified DFA3D CUDA
extension compiled from /root/reference by oracle/build_ref.py into oracle/_ref (travels to the GPU box).

Runs the reference autograd stitching of multi_scale_3ddeformable_attn_function.py:277-351 literally
(two forward calls, two backward calls, uv-grad add) on the reference extension."""
import pytest
import torch

import sgcdet_b200
from oracle import build_ref, dfa3d_ref

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-3, 1e-4


@pytest.fixture(scope='module')
def ref_ext():
    ext = build_ref.load()
    if ext is None:
        pytest.skip('oracle/_ref/dfa3d_ref_ext.so not built (needs /root/reference at build time)')
    return ext


def reference_fwd_bwd(ext, value, dist, s3, lsi, loc, attn, gout, step=64):
    """F3D:277-351 restated call-for-call on the reference extension."""
    ds = ext.ms_depth_score_sample_forward(dist, s3, lsi, loc, im2col_step=step)
    out = ext.wms_deform_attn_forward(value, s3[..., :2].contiguous(), lsi, loc[..., :2].contiguous(), attn, ds,
                                      im2col_step=step)
    g_value = torch.zeros_like(value)
    g_loc2 = torch.zeros([*loc.shape[:-1], 2], dtype=loc.dtype, device=loc.device)
    g_attn = torch.zeros_like(attn)
    g_ds = torch.zeros_like(ds)
    ext.wms_deform_attn_backward(value, s3[..., :2].contiguous(), lsi, loc[..., :2].contiguous(), attn, ds,
                                 gout.contiguous(), g_value, g_loc2, g_attn, g_ds, im2col_step=step)
    g_dist = torch.zeros_like(dist)
    g_loc = torch.zeros_like(loc)
    ext.ms_depth_score_sample_backward(dist, s3, lsi, loc, g_ds.contiguous(), g_dist, g_loc, im2col_step=step)
    g_loc[..., :2] = g_loc[..., :2] + g_loc2
    return dict(out=out, ds=ds, g_value=g_value, g_dist=g_dist, g_loc=g_loc, g_attn=g_attn, g_ds=g_ds)


def make_case(B, Q, M, Cm, D, shapes, P, seed, spread=1.3):
    g = torch.Generator().manual_seed(seed)
    s3 = torch.tensor([[h, w, D] for h, w in shapes], dtype=torch.long)
    sizes = s3[:, 0] * s3[:, 1]
    lsi = torch.cat([sizes.new_zeros(1), sizes.cumsum(0)[:-1]])
    S, L = int(sizes.sum()), len(shapes)
    return dict(value=torch.randn(B, S, M, Cm, generator=g), dist=torch.randn(B, S, M, D, generator=g).softmax(-1),
                s3=s3, lsi=lsi, loc=(torch.rand(B, Q, M, L, P, 3, generator=g) - 0.5) * spread + 0.5,
                attn=torch.rand(B, Q, M, L, P, generator=g), gout=torch.randn(B, Q, M * Cm, generator=g))


CASES = {
    'sgcdet_stage2': (4, 60, 8, 32, 12, [(14, 20)], 4),
    'sgcdet_stage1': (4, 60, 1, 256, 12, [(14, 20)], 1),
    'sgcdet_large': (2, 50, 8, 16, 12, [(7, 10)], 4),
    'unittest_family': (2, 64, 8, 32, 28, [(12, 20), (6, 10), (3, 5), (2, 3)], 8),
}


@pytest.mark.parametrize('name', list(CASES))
def test_oracle_matches_reference_kernels(ref_ext, name):
    c = make_case(*CASES[name], seed=3)
    cu = {k: v.cuda() for k, v in c.items()}
    ref = reference_fwd_bwd(ref_ext, cu['value'], cu['dist'], cu['s3'], cu['lsi'], cu['loc'], cu['attn'], cu['gout'])
    out_o, ds_o = dfa3d_ref.dfa3d_forward(c['value'], c['dist'], c['s3'], c['lsi'], c['loc'], c['attn'])
    torch.testing.assert_close(ref['ds'].cpu(), ds_o, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(ref['out'].cpu(), out_o, rtol=RTOL, atol=ATOL)
    gv, gd, gl, ga = dfa3d_ref.dfa3d_backward(c['value'], c['dist'], c['s3'], c['lsi'], c['loc'], c['attn'], c['gout'])
    torch.testing.assert_close(ref['g_value'].cpu(), gv, rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(ref['g_dist'].cpu(), gd, rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(ref['g_loc'].cpu(), gl, rtol=RTOL, atol=2e-3)
    torch.testing.assert_close(ref['g_attn'].cpu(), ga, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize('name', list(CASES))
def test_product_matches_reference_kernels(ref_ext, cuda_lib, name):
    sgcdet_b200.install_dropin()
    from dfa3D.ops import MultiScale3DDeformableAttnFunction
    c = make_case(*CASES[name], seed=9)
    cu = {k: v.cuda() for k, v in c.items()}
    ref = reference_fwd_bwd(ref_ext, cu['value'], cu['dist'], cu['s3'], cu['lsi'], cu['loc'], cu['attn'], cu['gout'])
    v, d, l, a = (cu[k].clone().requires_grad_(True) for k in ('value', 'dist', 'loc', 'attn'))
    out, ds = MultiScale3DDeformableAttnFunction.apply(v, d, cu['s3'], cu['lsi'], l, a, 64)
    out.backward(cu['gout'])
    torch.testing.assert_close(ds, ref['ds'], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(out, ref['out'], rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(v.grad, ref['g_value'], rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(d.grad, ref['g_dist'], rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(l.grad, ref['g_loc'], rtol=RTOL, atol=2e-3)
    torch.testing.assert_close(a.grad, ref['g_attn'], rtol=RTOL, atol=ATOL)


def test_two_stage_product_exports_match_reference(ref_ext, cuda_lib):
    """The four `_ext` functions one by one, including grad_depth_score and the '=' (not '+=') semantics of
    grad_sampling_loc / grad_attn_weight (WMSK:400-406)."""
    sgcdet_b200.install_dropin()
    from dfa3D import ext_loader
    ext = ext_loader.load_ext('_ext', ['wms_deform_attn_backward', 'wms_deform_attn_forward',
                                       'ms_depth_score_sample_forward', 'ms_depth_score_sample_backward'])
    c = make_case(*CASES['sgcdet_stage2'], seed=21)
    cu = {k: v.cuda() for k, v in c.items()}
    ref = reference_fwd_bwd(ref_ext, cu['value'], cu['dist'], cu['s3'], cu['lsi'], cu['loc'], cu['attn'], cu['gout'])
    got = reference_fwd_bwd(ext, cu['value'], cu['dist'], cu['s3'], cu['lsi'], cu['loc'], cu['attn'], cu['gout'])
    for k in ref:
        torch.testing.assert_close(got[k], ref[k], rtol=RTOL, atol=2e-3 if k == 'g_loc' else ATOL, msg=lambda m: f'{k}: {m}')


@pytest.mark.parametrize('V', [9, 20])
def test_product_matches_reference_kernels_end_to_end(ref_ext, cuda_lib, V):
    """Whole view transform: the reference's own DFA3D kernels under the restated reference glue (oracle/gpu_ref.py:
    padded per-view rebatch, torch MHA, F.interpolate, torch.topk) vs the fused sgcdet_b200 path, same weights and scene.
    The product is teacher-forced with the reference's selection."""
    from oracle import gpu_ref
    from sgcdet_b200 import plugin, synthetic as syn
    cfg = syn.CONFIGS['tiny']
    sc = syn.make_scene(cfg, V, shift_origin=True).to('cuda')
    sd = syn.make_state_dict(cfg)
    sdg = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad():
        vol_r, valid_r, occ_r = gpu_ref.head_forward_gpu(sdg, sc.mlvl_feats, sc.img_meta, sc.mlvl_dpt_dists, cfg, training=False)
    head = plugin.build_voxel_head(cfg)
    head.load_state_dict(sd)
    head = head.cuda().eval()
    n_mid = int(torch.tensor(cfg.n_voxels_list[1]).prod())
    n_fine = int(torch.tensor(cfg.n_voxels_list[2]).prod())
    # reference selections: finest from `valid`, middle from the top-k of its own occupancy
    sel_fine = torch.nonzero(valid_r.view(-1)).view(-1).to(torch.int32)
    occ_mid = occ_r[0, n_fine:n_fine + n_mid]
    sel_mid = torch.sort(torch.topk(occ_mid, cfg.topk_list[0]).indices).values.to(torch.int32)
    with torch.no_grad():
        vol, valid, occ = head(sc.mlvl_feats, sc.img_meta, sc.mlvl_dpt_dists, forced_selection=[None, sel_mid, sel_fine])
    assert torch.equal(valid, valid_r)
    torch.testing.assert_close(occ, occ_r, rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(vol, vol_r, rtol=RTOL, atol=ATOL)


def test_two_stage_exports_fp64_match_reference(ref_ext, cuda_lib):
    """The reference dispatches the four `_ext` functions over float AND double (wms_deform_attn_cuda.cu:267,344,
    ms_depth_score_sample_cuda.cu:95,153): the drop-in's fp64 instantiation (csrc/dfa3d_op_f64.cu) against the reference's own
    kernels on double inputs, and against the fp32 path."""
    sgcdet_b200.install_dropin()
    from dfa3D import ext_loader
    ext = ext_loader.load_ext('_ext', ['wms_deform_attn_backward', 'wms_deform_attn_forward',
                                       'ms_depth_score_sample_forward', 'ms_depth_score_sample_backward'])
    for name in ('sgcdet_stage2', 'unittest_family'):
        c = make_case(*CASES[name], seed=33)
        cd = {k: (v.double().cuda() if v.is_floating_point() else v.cuda()) for k, v in c.items()}
        ref = reference_fwd_bwd(ref_ext, cd['value'], cd['dist'], cd['s3'], cd['lsi'], cd['loc'], cd['attn'], cd['gout'])
        got = reference_fwd_bwd(ext, cd['value'], cd['dist'], cd['s3'], cd['lsi'], cd['loc'], cd['attn'], cd['gout'])
        for k in ref:
            assert got[k].dtype == torch.float64
            torch.testing.assert_close(got[k], ref[k], rtol=1e-9, atol=1e-10, msg=lambda m: f'{name}/{k}: {m}')
        cf = {k: v.cuda() for k, v in c.items()}
        g32 = reference_fwd_bwd(ext, cf['value'], cf['dist'], cf['s3'], cf['lsi'], cf['loc'], cf['attn'], cf['gout'])
        torch.testing.assert_close(g32['out'].double(), got['out'], rtol=RTOL, atol=ATOL)
    # mixed dtypes are refused like a failed dispatch
    with pytest.raises(RuntimeError):
        ext.ms_depth_score_sample_forward(cd['dist'], cd['s3'], cd['lsi'], cd['loc'].float(), im2col_step=64)
