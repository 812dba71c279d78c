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
