"""Per-step weight preparation (one launch) and the per-voxel row kernels (csrc/sgc_rowops.cu).
Bit-exact against the single-matrix split / pack kernels; LayerNorm backward against torch autograd
(rtol 1e-3 / atol 1e-4 relative to the tensor scale)."""
import math

import pytest
import torch

from sgcdet_b200 import functional as SF

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('C,N,F', [(256, 384, 512), (128, 256, 256)])
def test_level_weights_one_launch_matches_single_kernels(cuda_lib, C, N, F):
    g = torch.Generator().manual_seed(C + N)
    dev = 'cuda'
    wcat = torch.randn(N, C, generator=g).to(dev)
    w_out, wo = torch.randn(C, C, generator=g).to(dev), torch.randn(C, C, generator=g).to(dev)
    in_w = torch.randn(3 * C, C, generator=g).to(dev)
    w1, w2 = torch.randn(F, C, generator=g).to(dev), torch.randn(C, F, generator=g).to(dev)
    lw = SF.LevelWeights(wcat, w_out, in_w, wo, w1, w2)
    dh = C // 8
    scale = 1.0 / math.sqrt(dh)
    wq, wk, wv = in_w[:C], in_w[C:2 * C] * scale, in_w[2 * C:]
    ref = dict(wcat=SF.split_cols(wcat, 1), wcat_t=SF.split_cols(wcat.t(), 1),
               wpack=SF.pack_weight_tc(wcat), wpack_t=SF.pack_weight_tc(wcat.t().contiguous()),
               w_out=SF.split_cols(w_out, 1), w_out_t=SF.split_cols(w_out.t(), 1),
               wq=SF.split_cols(wq, 1), wq_t=SF.split_cols(wq.t(), 1),
               wo=SF.split_cols(wo, 1), wo_t=SF.split_cols(wo.t(), 1),
               wk_rows=SF.split_rows(wk, dh, 1), wk_cols=SF.split_cols(wk, 1),
               wv_rows=SF.split_rows(wv, dh, 1), wv_cols=SF.split_cols(wv, 1),
               w1=SF.split_cols(w1, 1), w1_t=SF.split_cols(w1.t(), 1),
               w2=SF.split_cols(w2, 1), w2_t=SF.split_cols(w2.t(), 1))
    torch.cuda.synchronize()
    for k, r in ref.items():
        got = getattr(lw, k)
        assert got.shape == r.shape, k
        assert torch.equal(got.view(torch.int16), r.view(torch.int16)), k


@pytest.mark.parametrize('R,C', [(1, 256), (37, 128), (400, 256), (6400, 256), (51200, 128)])
def test_layernorm_rows_backward_matches_torch(cuda_lib, R, C):
    g = torch.Generator().manual_seed(R + C)
    x = (torch.randn(R, C, generator=g) * 3 + 0.5).cuda()
    gy = torch.randn(R, C, generator=g).cuda()
    gamma, beta = torch.randn(C, generator=g).cuda(), torch.randn(C, generator=g).cuda()
    a = [t.clone().double().requires_grad_(True) for t in (x, gamma, beta)]
    torch.nn.functional.layer_norm(a[0], (C,), a[1], a[2], 1e-5).backward(gy.double())
    b = [t.clone().requires_grad_(True) for t in (x, gamma, beta)]
    y = SF.LayerNormRows.apply(b[0], b[1], b[2], 1e-5, None)
    y.backward(gy)
    torch.cuda.synchronize()
    ref_y = torch.nn.functional.layer_norm(x, (C,), gamma, beta, 1e-5)
    assert torch.equal(y, ref_y)
    for got, ref, name in zip(b, a, ('x', 'gamma', 'beta')):
        scale = max(ref.grad.abs().max().item(), 1e-6)
        err = (got.grad.double() - ref.grad).abs().max().item()
        assert err <= 1e-4 + 1e-3 * scale, (name, err, scale)


def test_detached_weight_stream_gradients_match(cuda_lib):
    """Weight gradients produced on the weight stream (never joined into the calling stream) equal the joined ones."""
    g = torch.Generator().manual_seed(7)
    Q, C, F = 800, 256, 512
    x0 = torch.randn(Q, C, generator=g).cuda()
    w0, b0 = (torch.randn(F, C, generator=g) / 16).cuda(), torch.randn(F, generator=g).cuda()
    gy = torch.randn(Q, F, generator=g).cuda()
    res = []
    for detached in (False, True):
        x, w, b = (t.clone().requires_grad_(True) for t in (x0, w0, b0))
        ws = torch.cuda.Stream() if detached else None
        if detached:
            with torch.cuda.stream(ws):
                wa, ba = SF.OnStream.apply(w, b)
        else:
            wa, ba = w, b
        for _ in range(3):  # accumulate: exercises AccumulateGrad's in-place path on the weight stream
            y = SF.Linear3.apply(x, wa, ba, None, None, ws)
            y.backward(gy, retain_graph=True)
        torch.cuda.synchronize()
        res.append((x.grad.clone(), w.grad.clone(), b.grad.clone()))
    for a, b_ in zip(*res):
        assert torch.equal(a, b_)


def _split_ref(x, heads=0):
    """bf16x3 pattern-0 image computed with torch ops."""
    hi = x.bfloat16()
    lo = (x - hi.float()).bfloat16()
    if heads:
        R, N = x.shape
        dh = N // heads
        h3, l3 = hi.view(R, heads, dh), lo.view(R, heads, dh)
        return torch.cat([h3, l3, h3], dim=2).reshape(R * heads, 3 * dh)
    return torch.cat([hi, lo, hi], dim=1)


@pytest.mark.parametrize('R,N', [(5, 128), (400, 256), (801, 512), (6400, 256)])
def test_rowop_fwd_bwd_match_torch(cuda_lib, R, N):
    g = torch.Generator().manual_seed(R * 7 + N)
    dev = 'cuda'
    x = torch.randn(R, N, generator=g).to(dev)
    bias, gamma, beta = (torch.randn(N, generator=g).to(dev) for _ in range(3))
    mask = (torch.rand(R, N, generator=g) > 0.3).to(torch.uint8).to(dev)
    rowscale = (torch.rand(R, generator=g) > 0.2).float().to(dev)
    res = torch.randn(R, N, generator=g).to(dev)
    gy = torch.randn(R, N, generator=g).to(dev)
    ms = 1.0 / 0.7

    xa = x.clone().double().requires_grad_(True)
    ra = res.clone().double().requires_grad_(True)
    ga, ba = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    pre = torch.relu(xa + bias.double()) * mask.double() * ms * rowscale.double()[:, None] + ra
    ya = torch.nn.functional.layer_norm(pre, (N,), ga, ba, 1e-5)
    ya.backward(gy.double())

    y, ys, saved = SF.rowop_fwd(x, R, N, bias=bias, relu=True, mask=mask, mscale=ms, rowscale=rowscale, residual=res,
                                ln=(gamma, beta, 1e-5))
    torch.cuda.synchronize()
    assert (y.double() - ya).abs().max().item() < 1e-4
    assert torch.equal(ys.view(torch.int16), _split_ref(y).view(torch.int16))
    assert (saved[0].double() - pre).abs().max().item() < 1e-5
    # backward through LN, then mask / gate / rowscale (gate = relu output, as the FFN uses it)
    hdn = torch.relu(x + bias)
    gx, gs, gpre, partial = SF.rowop_bwd(gy, R, N, ln=(saved[0], saved[1], saved[2], gamma), mask=mask, mscale=ms,
                                         gate=hdn, gscale=1.0, rowscale=rowscale, want_gpre=True)
    gg, gb = SF._ln_params(partial, R, N)
    torch.cuda.synchronize()
    scale = max(xa.grad.abs().max().item(), 1e-6)
    assert (gpre.double() - ra.grad).abs().max().item() <= 1e-4 + 1e-3 * scale
    assert (gx.double() - xa.grad).abs().max().item() <= 1e-4 + 1e-3 * scale
    assert torch.equal(gs.view(torch.int16), _split_ref(gx).view(torch.int16))
    for got, ref in ((gg, ga.grad), (gb, ba.grad)):
        sc = max(ref.abs().max().item(), 1e-6)
        assert (got.double() - ref).abs().max().item() <= 1e-4 + 1e-3 * sc


def test_rowop_rowcount_equals_rowscale(cuda_lib):
    """``rowcount`` (per-voxel view counts) zeroes exactly the rows a 0/1 ``rowscale`` built from it would."""
    g = torch.Generator().manual_seed(9)
    R, N = 777, 256
    x, gy = torch.randn(R, N, generator=g).cuda(), torch.randn(R, N, generator=g).cuda()
    bias, gamma, beta = (torch.randn(N, generator=g).cuda() for _ in range(3))
    count = torch.randint(0, 3, (R,), generator=g).to(torch.int32).cuda()
    has = (count > 0).float()
    a = SF.rowop_fwd(x, R, N, bias=bias, rowscale=has, ln=(gamma, beta, 1e-5))
    b = SF.rowop_fwd(x, R, N, bias=bias, rowcount=count, ln=(gamma, beta, 1e-5))
    ga = SF.rowop_bwd(gy, R, N, ln=(a[2][0], a[2][1], a[2][2], gamma), rowscale=has)
    gb = SF.rowop_bwd(gy, R, N, ln=(b[2][0], b[2][1], b[2][2], gamma), rowcount=count)
    torch.cuda.synchronize()
    assert (count == 0).any()
    assert torch.equal(a[0], b[0]) and torch.equal(a[1].view(torch.int16), b[1].view(torch.int16))
    assert torch.equal(ga[0], gb[0]) and torch.equal(ga[3], gb[3])
    assert (gb[0][count == 0] == 0).all()


def test_rowop_head_layouts(cuda_lib):
    g = torch.Generator().manual_seed(3)
    R, N, H = 333, 256, 8
    dh = N // H
    xh = torch.randn(H, R, dh, generator=g).cuda()
    bias = torch.randn(N, generator=g).cuda()
    y, ys, _ = SF.rowop_fwd(xh, R, N, bias=bias, in_heads=H, split_heads=H)
    ref = xh.permute(1, 0, 2).reshape(R, N) + bias
    torch.cuda.synchronize()
    assert torch.equal(y, ref)
    assert torch.equal(ys.view(R * H, 3 * dh).view(torch.int16), _split_ref(ref, H).view(torch.int16))
    gx, gs, _, _ = SF.rowop_bwd(xh, R, N, in_heads=H)
    assert torch.equal(gx, xh.permute(1, 0, 2).reshape(R, N))
    assert torch.equal(gs.view(torch.int16), _split_ref(gx).view(torch.int16))


@pytest.mark.parametrize('train', [False, True])
def test_fused_layer_matches_unfused_path(cuda_lib, train, monkeypatch):
    """EncoderLayerRows (fused row kernels) against the op-by-op path on the tiny config: same outputs and gradients.
    In train mode both paths get the same dropout keep-masks."""
    from sgcdet_b200 import plugin, synthetic as syn
    cfg = syn.CONFIGS['tiny']
    sc = syn.make_scene(cfg, 6, shift_origin=True).to('cuda')
    res = []
    forced = None  # the second run is teacher-forced with the first run's selection (top-k ties flip at round-off)
    for fused in ('1', '0'):
        monkeypatch.setenv('SGC_FUSED_LAYER', fused)
        head = plugin.build_voxel_head(cfg)
        head.load_state_dict(syn.make_state_dict(cfg))
        head = head.cuda().train(train)
        if train:
            for m in head.modules():
                if isinstance(m, torch.nn.Dropout):
                    m.p = 0.0  # masks are compared through the eval-equivalent path: dropout off in both
        feats = [f.clone().requires_grad_(True) for f in sc.mlvl_feats[:3]]
        dists = [d.clone().requires_grad_(True) for d in sc.mlvl_dpt_dists[:3]]
        vol, valid, occ, inters = head(feats, sc.img_meta, dists, forced_selection=forced, return_intermediates=True)
        if forced is None:
            forced = [it['sel'] for it in inters]
        loss = (vol * sc.grad_volume).sum() + head.occ_loss(occ, None, sc.geo_occ)['loss_occ']
        loss.backward()
        torch.cuda.synchronize()
        res.append((vol.detach(), occ.detach(), [f.grad for f in feats], [d.grad for d in dists],
                    {n: p.grad for n, p in head.named_parameters() if p.grad is not None}))
    a, b = res
    assert (a[0] - b[0]).abs().max().item() <= 1e-4 + 1e-3 * b[0].abs().max().item()
    assert (a[1] - b[1]).abs().max().item() <= 1e-5

    def close(x, y, name):
        # ReLU gates at round-off flip between the two paths, so single entries may differ: norm-wise check plus a
        # loose bound on the worst entry
        fro = ((x - y).norm() / y.norm().clamp_min(1e-12)).item()
        worst = ((x - y).abs().max() / y.abs().max().clamp_min(1e-12)).item()
        assert fro < 1e-2 and worst < 5e-2, (name, fro, worst)
    for i, (ga, gb) in enumerate(zip(a[2] + a[3], b[2] + b[3])):
        close(ga, gb, f'input{i}')
    assert a[4].keys() == b[4].keys()
    for n in a[4]:
        close(a[4][n], b[4][n], n)


def test_fold_weights_matches_cat_and_unfolds_gradients(cuda_lib):
    """sgc_fold_wcat / sgc_unfold_wcat_grad (the folded projection weights of MSDeformableAttention3D_DFA3D) against the torch
    cat / view formulation (the CPU branch of ``folded_weights``): forward bit-exact, gradients bit-exact copies."""
    import torch
    from sgcdet_b200 import plugin
    for C in (128, 256):
        da_cpu = plugin.MSDeformableAttention3D_DFA3D(embed_dims=C, num_heads=8, num_levels=1, num_points=4)
        g = torch.Generator().manual_seed(C)
        with torch.no_grad():
            for p in da_cpu.parameters():
                p.copy_(torch.randn(p.shape, generator=g))
        da = plugin.MSDeformableAttention3D_DFA3D(embed_dims=C, num_heads=8, num_levels=1, num_points=4)
        da.load_state_dict(da_cpu.state_dict())
        da = da.cuda()
        wcat_r, vb_r, gb_r = da_cpu.folded_weights()
        wcat, vb, gb = da.folded_weights()
        assert torch.equal(wcat.cpu(), wcat_r) and torch.equal(gb.cpu(), gb_r) and torch.equal(vb.cpu(), vb_r)
        gw = torch.randn(wcat.shape, generator=g)
        gg = torch.randn(gb.shape, generator=g)
        (wcat_r * gw).sum().add((gb_r * gg).sum()).backward()
        (wcat * gw.cuda()).sum().add((gb * gg.cuda()).sum()).backward()
        for (k, p), (_, q) in zip(da.named_parameters(), da_cpu.named_parameters()):
            if k == 'value_proj.bias':
                continue
            assert p.grad is not None and p.grad.is_contiguous(), k
            assert torch.equal(p.grad.cpu(), q.grad), k
