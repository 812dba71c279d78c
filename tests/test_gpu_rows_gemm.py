"""Voxel-count GEMMs on the own tcgen05 kernel (csrc/sgc_rows_gemm_tc.cu) vs fp64 references of the same products.
Tolerance: rtol 1e-3 / atol 1e-4 relative to the output scale (the bf16 hi/lo split gives ~1e-5)."""
import math

import pytest
import torch

from sgcdet_b200 import functional as SF
from sgcdet_b200._lib import call, ptr, stream

pytestmark = pytest.mark.gpu


def pack(w):
    N, K = w.shape
    out = torch.empty(2 * N * K, device=w.device, dtype=torch.bfloat16)
    call('sgc_pack_weight_tc', ptr(w.contiguous()), N, K, ptr(out), stream())
    return out


def check(got, ref):
    assert torch.isfinite(got).all()
    scale = ref.abs().max().item()
    err = (got.double() - ref).abs().max().item() / scale
    assert err < 1e-4, err
    torch.testing.assert_close(got.double() / scale, ref / scale, rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize('R,K,N,n_cta,bias', [
    (6400, 256, 256, 0, True),     # output_proj / attention in-projection at the finest level
    (6400, 256, 512, 0, True),     # FFN layer 1
    (6400, 512, 256, 0, False),    # FFN layer 2 (16 k-slabs)
    (6400, 256, 256, 256, False),  # one column part per row tile
    (6400, 256, 256, 32, True),    # eight column parts per row tile
    (800, 256, 256, 0, True),      # 7 row tiles, the last one 32 rows
    (400, 256, 512, 64, False),
    (37, 256, 256, 0, True),       # fewer rows than one tile
    (1, 128, 128, 0, True),
    (3200, 128, 256, 128, True),   # "-L" shapes
    (51200, 128, 128, 0, False),   # more work items than SMs: every pipeline wraps several times
    (20000, 32, 64, 64, True),     # a single k-slab per work item
])
def test_rows_linear_matches_fp64(cuda_lib, R, K, N, n_cta, bias):
    g = torch.Generator().manual_seed(R + 7 * K + 13 * N + n_cta)
    x = torch.randn(R, K, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    b = torch.randn(N, generator=g).cuda() if bias else None
    y = SF.rows_linear(x, pack(w), N, b, n_cta)
    torch.cuda.synchronize()
    ref = x.double() @ w.double().t()
    if bias:
        ref = ref + b.double()
    check(y, ref)


def test_rows_gemm_clips_rows_and_leaves_the_rest_untouched(cuda_lib):
    """Rows >= R of the output allocation are never written (the TMA store clips), padding columns neither."""
    g = torch.Generator().manual_seed(3)
    R, K, N, ld = 200, 256, 128, 192
    x = torch.randn(R, K, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) / 16).cuda()
    y = torch.full((R + 100, ld), 7.0, device='cuda')
    call('sgc_rows_gemm_tc', ptr(x), K, 0, R, K, 1, ptr(pack(w)), N, 0, 0, None, 0, N, ptr(y), ld, 0, 0, stream())
    torch.cuda.synchronize()
    assert (y[R:] == 7.0).all() and (y[:, N:] == 7.0).all()
    check(y[:R, :N], x.double() @ w.double().t())


@pytest.mark.parametrize('R,C,N', [(6400, 256, 256), (800, 256, 256), (77, 256, 256)])
def test_rows_heads_in_matches_fp64(cuda_lib, R, C, N):
    """y[h] = x[:, h*dh:(h+1)*dh] @ W_h with the per-head packed transposes produced by sgc_prepare_weights."""
    H = 8
    dh = C // H
    g = torch.Generator().manual_seed(R + C)
    x = torch.randn(R, C, generator=g).cuda()
    wk = (torch.randn(C, N, generator=g) / dh ** 0.5).cuda()      # rows h*dh.. are W_h [dh, N]
    scale = 1.0 / math.sqrt(dh)
    j = SF._WeightJobs(x.device)
    p = j.pack_heads_t(wk, H, scale)
    j.launch()
    y = SF.rows_heads_in(x, p, N, H)
    torch.cuda.synchronize()
    ref = torch.einsum('rhd,hdn->hrn', x.double().view(R, H, dh), wk.double().view(H, dh, N)) * scale
    check(y, ref)


@pytest.mark.parametrize('R,C', [(6400, 256), (800, 256), (77, 256)])
def test_rows_heads_out_matches_fp64(cuda_lib, R, C):
    """y[:, h*dh:(h+1)*dh] = x[h] @ W[h*dh:(h+1)*dh]^T + bias, written straight into the [R,C] layout."""
    H = 8
    dh = C // H
    g = torch.Generator().manual_seed(R * 3 + C)
    x = torch.randn(H, R, C, generator=g).cuda()
    wv = (torch.randn(C, C, generator=g) / C ** 0.5).cuda()
    b = torch.randn(C, generator=g).cuda()
    y = SF.rows_heads_out(x, pack(wv), dh, b)
    torch.cuda.synchronize()
    ref = torch.einsum('hrk,hdk->rhd', x.double(), wv.double().view(H, dh, C)).reshape(R, C) + b.double()
    check(y, ref)


def test_level_weights_packs_match_single_kernel(cuda_lib):
    """The packed operands LevelWeights prepares in one launch are bit-identical to sgc_pack_weight_tc of the same
    (scaled / transposed) matrices."""
    C, N, F = 256, 384, 512
    g = torch.Generator().manual_seed(11)
    dev = 'cuda'
    wcat = torch.randn(N, C, generator=g).to(dev)
    w_out, wo = torch.randn(C, C, generator=g).to(dev), torch.randn(C, C, generator=g).to(dev)
    in_w = torch.randn(3 * C, C, generator=g).to(dev)
    w1, w2 = torch.randn(F, C, generator=g).to(dev), torch.randn(C, F, generator=g).to(dev)
    lw = SF.LevelWeights(wcat, w_out, in_w, wo, w1, w2)
    dh = C // 8
    scale = 1.0 / math.sqrt(dh)
    wq, wk, wv = in_w[:C], in_w[C:2 * C], in_w[2 * C:]
    ref = dict(p_w_out=pack(w_out), p_w_out_t=pack(w_out.t()), p_wq=pack(wq), p_wq_t=pack(wq.t()), p_wo=pack(wo),
               p_wo_t=pack(wo.t()), p_w1=pack(w1), p_w1_t=pack(w1.t()), p_w2=pack(w2), p_w2_t=pack(w2.t()),
               p_wk=pack(wk * scale), p_wv=pack(wv),
               p_wk_ht=torch.cat([pack((wk[h * dh:(h + 1) * dh] * scale).t()) for h in range(8)]),
               p_wv_ht=torch.cat([pack(wv[h * dh:(h + 1) * dh].t()) for h in range(8)]))
    torch.cuda.synchronize()
    for k, r in ref.items():
        got = getattr(lw, k)
        assert got.shape == r.shape, k
        assert torch.equal(got.view(torch.int16), r.view(torch.int16)), k


@pytest.mark.parametrize('R,N,K', [
    (6400, 256, 256),    # output_proj / in-projection weight gradients at the finest level
    (6400, 512, 256),    # FFN layer 1 (four 128-row tiles of the gradient)
    (6400, 256, 512),    # FFN layer 2 (two column parts)
    (800, 256, 256),
    (400, 256, 512),     # a single k-chunk
    (37, 128, 128),      # fewer rows than two slabs
    (51200, 128, 256),   # "-L" finest level
])
def test_linear_grads_tc_matches_fp64(cuda_lib, R, N, K):
    g = torch.Generator().manual_seed(R + 3 * N + 5 * K)
    gy = torch.randn(R, N, generator=g).cuda()
    x = torch.randn(R, K, generator=g).cuda()
    gw, gb = SF.linear_grads_tc(gy, x)
    torch.cuda.synchronize()
    check(gw, gy.double().t() @ x.double())
    check(gb, gy.double().sum(0))


def test_linear_grads_tc_is_deterministic_and_writes_in_place(cuda_lib):
    g = torch.Generator().manual_seed(21)
    R, C = 3000, 256
    gy, x = torch.randn(R, C, generator=g).cuda(), torch.randn(R, C, generator=g).cuda()
    big_w = torch.full((3 * C, C), 5.0, device='cuda')
    big_b = torch.full((3 * C,), 5.0, device='cuda')
    SF.linear_grads_tc(gy, x, big_w[C:2 * C], big_b[C:2 * C])
    a, _ = SF.linear_grads_tc(gy, x)
    b, _ = SF.linear_grads_tc(gy, x)
    torch.cuda.synchronize()
    assert torch.equal(a, b) and torch.equal(a, big_w[C:2 * C])
    assert (big_w[:C] == 5.0).all() and (big_w[2 * C:] == 5.0).all()
    assert (big_b[:C] == 5.0).all() and (big_b[2 * C:] == 5.0).all()


@pytest.mark.parametrize('R', [6400, 800, 77])
def test_rows_wgrad_heads_matches_fp64(cuda_lib, R):
    """Per-head key / value weight gradients: out[h*dh + d, c] = scale * sum_r a[h][r, c] * b[r, h*dh + d], written
    transposed into a row slice of in_proj_weight's gradient; bias gradient = column sums of b."""
    H, C = 8, 256
    dh = C // H
    g = torch.Generator().manual_seed(R)
    a = torch.randn(H, R, C, generator=g).cuda()
    b = torch.randn(R, C, generator=g).cuda()
    gw = torch.full((3 * C, C), 9.0, device='cuda')
    gb = torch.full((3 * C,), 9.0, device='cuda')
    scale = 0.37
    SF.rows_wgrad(a, b, C, dh, R, gw[2 * C:], (dh * C, 1, C), B=H, lda=C, batch_a=R * C, ldb=C, batch_b=dh, scale=scale,
                  bias_out=gb[2 * C:], bias_from=2)
    torch.cuda.synchronize()
    ref = torch.einsum('hrc,rhd->hdc', a.double(), b.double().view(R, H, dh)).reshape(C, C) * scale
    check(gw[2 * C:], ref)
    check(gb[2 * C:], b.double().sum(0))
    assert (gw[:2 * C] == 9.0).all() and (gb[:2 * C] == 9.0).all()


@pytest.mark.parametrize('R,C', [(6400, 256), (800, 256), (400, 256), (37, 256), (51200, 128), (3200, 128)])
def test_wgrad_group_matches_fp64(cuda_lib, R, C):
    """All seven weight-gradient products of an encoder layer as ONE grouped launch (sgc_rows_wgrad_group_tc): plain Linear
    layers, the per-head key / value products written transposed into in_proj_weight's gradient (32- and 16-wide heads),
    bias gradients from either operand, scale; untouched rows of the destination stay untouched; deterministic."""
    H, F = 8, 2 * C
    dh = C // H
    g = torch.Generator().manual_seed(R + C)
    rnd = lambda *s: torch.randn(*s, generator=g).cuda()
    gf, hdn, gh, x1, gout, o2 = rnd(R, C), rnd(R, F), rnd(R, F), rnd(R, C), rnd(R, C), rnd(R, C)
    t, go2, gqt, qv, gqv, gg_, gmean_in, mean = rnd(H, R, C), rnd(R, C), rnd(H, R, C), rnd(R, C), rnd(R, C), rnd(R, C), rnd(R, C), rnd(R, C)
    scale = 1.0 / math.sqrt(dh)

    def run():
        G = SF.WgradGroup(R, 'cuda')
        gw_in = torch.full((3 * C, C), 9.0, device='cuda')
        gb_in = torch.full((3 * C,), 9.0, device='cuda')
        w2 = G.linear(gf, hdn)
        w1 = G.linear(gh, x1)
        wo = G.linear(gout, o2)
        G.add(t, go2, C, dh, gw_in[2 * C:], (dh * C, 1, C), B=H, lda=C, batch_a=R * C, ldb=C, batch_b=dh,
              bias_out=gb_in[2 * C:], bias_from=2)
        G.add(gqt, qv, C, dh, gw_in[C:2 * C], (dh * C, 1, C), B=H, lda=C, batch_a=R * C, ldb=C, batch_b=dh, scale=scale)
        G.linear(gqv, gg_, gw_in[:C], gb_in[:C])
        wout = G.linear(gmean_in, mean)
        G.launch()
        torch.cuda.synchronize()
        return w2, w1, wo, wout, gw_in, gb_in

    (g_w2, g_b2), (g_w1, g_b1), (g_wo, g_bo), (g_wout, g_bout), gw_in, gb_in = run()
    d = lambda x: x.double()
    check(g_w2, d(gf).t() @ d(hdn)); check(g_b2, d(gf).sum(0))
    check(g_w1, d(gh).t() @ d(x1)); check(g_b1, d(gh).sum(0))
    check(g_wo, d(gout).t() @ d(o2)); check(g_bo, d(gout).sum(0))
    check(g_wout, d(gmean_in).t() @ d(mean)); check(g_bout, d(gmean_in).sum(0))
    check(gw_in[2 * C:], torch.einsum('hrc,rhd->hdc', d(t), d(go2).view(R, H, dh)).reshape(C, C))
    check(gb_in[2 * C:], d(go2).sum(0))
    check(gw_in[C:2 * C], torch.einsum('hrc,rhd->hdc', d(gqt), d(qv).view(R, H, dh)).reshape(C, C) * scale)
    assert (gb_in[C:2 * C] == 9.0).all()          # no bias job for the key projection: left untouched
    check(gw_in[:C], d(gqv).t() @ d(gg_)); check(gb_in[:C], d(gqv).sum(0))
    again = run()
    assert torch.equal(again[4], gw_in) and torch.equal(again[0][0], g_w2) and torch.equal(again[3][1], g_bout)


@pytest.mark.parametrize('R,C', [(3200, 128), (51200, 128), (77, 128), (400, 256)])
def test_narrow_head_products_match_fp64(cuda_lib, R, C):
    """Per-head key / value products for heads narrower than a k-slab (16 wide at C = 128; also valid for 32): the head's
    weights zero-extended over all C columns (a_mode 1: one shared activation matrix) and the per-head output products as one
    K-concatenated GEMM (a_mode 2), as LevelWeights packs them."""
    H = 8
    dh = C // H
    g = torch.Generator().manual_seed(R + C)
    w_out, wo = torch.randn(C, C, generator=g).cuda(), torch.randn(C, C, generator=g).cuda()
    in_w = (torch.randn(3 * C, C, generator=g) / C ** 0.5).cuda()
    w1, w2 = torch.randn(2 * C, C, generator=g).cuda(), torch.randn(C, 2 * C, generator=g).cuda()
    wcat = torch.randn(C + 128, C, generator=g).cuda()
    lw = SF.LevelWeights(wcat, w_out, in_w, wo, w1, w2)
    if lw.heads_tc:
        pytest.skip('32-wide heads use the per-head tensor maps (heads_tc)')
    scale = 1.0 / math.sqrt(dh)
    wk, wv = in_w[C:2 * C].double(), in_w[2 * C:].double()
    qv = torch.randn(R, C, generator=g).cuda()
    t = torch.randn(H, R, C, generator=g).cuda()
    bv = torch.randn(C, generator=g).cuda()
    qt = SF.rows_heads_in_exp(qv, lw.p_wk_in, H)
    gt = SF.rows_heads_in_exp(qv, lw.p_wv_in, H)
    o2 = SF.rows_heads_out_exp(t, lw.p_wv_out, bv)
    gqv = SF.rows_heads_out_exp(t, lw.p_wk_out)
    torch.cuda.synchronize()
    check(qt, torch.einsum('rhd,hdc->hrc', qv.double().view(R, H, dh), wk.view(H, dh, C)) * scale)
    check(gt, torch.einsum('rhd,hdc->hrc', qv.double().view(R, H, dh), wv.view(H, dh, C)))
    check(o2, torch.einsum('hrk,hdk->rhd', t.double(), wv.view(H, dh, C)).reshape(R, C) + bv.double())
    check(gqv, torch.einsum('hrk,hdk->rhd', t.double(), wk.view(H, dh, C)).reshape(R, C) * scale)
