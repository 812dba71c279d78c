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
               p_w_out=SF.pack_weight_tc(w_out), p_w_out_t=SF.pack_weight_tc(w_out.t().contiguous()),
               p_w1=SF.pack_weight_tc(w1), p_w2_t=SF.pack_weight_tc(w2.t().contiguous()))
    torch.cuda.synchronize()
    for k, r in ref.items():
        got = getattr(lw, k)
        assert got.shape == r.shape, k
        assert torch.equal(got.view(torch.int16), r.view(torch.int16)), k


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


def test_dropout_masks_match_the_philox_oracle_and_advance_per_launch(cuda_lib):
    """sgc_dropout_masks: bit-exact against oracle/philox_ref.py for every job of a launch, ragged tails included; the step
    counter on the device advances with every launch (so a CUDA-graph replay draws fresh masks)."""
    import numpy as np
    from oracle.philox_ref import keep_mask
    rng = SF.DropoutMasks('cuda', seed=1234)
    specs = [(37, 16, 0.0), (400, 256, 0.1), (401, 33, 0.25), (6400, 512, 0.1)]
    for step in range(3):
        masks = rng.draw(specs)
        torch.cuda.synchronize()
        assert masks[0] is None
        job = 0
        for (r, w, p), m in zip(specs[1:], masks[1:]):
            assert m.shape == (r, w) and m.dtype == torch.uint8
            ref = keep_mask(r * w, 1.0 - p, rng.seed, job, step).reshape(r, w)
            assert np.array_equal(m.cpu().numpy(), ref), (step, job)
            job += 1
        assert abs(masks[3].float().mean().item() - 0.9) < 1e-3
    assert int(rng.state[0]) == 3 and int(rng.state[1]) == 0
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        rng.draw(specs[1:2])           # warm-up outside the capture
        with torch.cuda.graph(g, stream=side):
            m = rng.draw(specs[1:2])[0]
    torch.cuda.current_stream().wait_stream(side)
    step0 = int(rng.state[0])
    seen = []
    for k in range(2):
        g.replay()
        torch.cuda.synchronize()
        assert np.array_equal(m.cpu().numpy(), keep_mask(400 * 256, 0.9, rng.seed, 0, step0 + k).reshape(400, 256))
        seen.append(m.clone())
    assert not torch.equal(seen[0], seen[1])
