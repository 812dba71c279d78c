"""Times the tcgen05 projection against the split + library-GEMM path at one level's shape."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from sgcdet_b200 import functional as SF
from sgcdet_b200._lib import call, ptr, stream

V, C, H0, W0, h, N = 40, 256, 60, 80, 59, 384
if len(sys.argv) > 1:
    V, C, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
S = h * W0
feat = torch.randn(V, C, H0, W0, device='cuda')
wcat = torch.randn(N, C, device='cuda') / C ** 0.5
wpack = torch.empty(2 * N * C, device='cuda', dtype=torch.bfloat16)
vg = torch.empty(V, S, N, device='cuda')


def tc():
    call('sgc_pack_weight_tc', ptr(wcat), N, C, ptr(wpack), stream())
    call('sgc_project_tc_fwd', ptr(feat), C * H0 * W0, H0 * W0, V, C, S, ptr(wpack), N, ptr(vg), stream())


def lib():
    acat = torch.empty(V, 3 * C, S, device='cuda', dtype=torch.bfloat16)
    call('sgc_split_bf16x3', ptr(feat), V * C, S, H0 * W0, C, 0, ptr(acat), stream())
    bcat = SF.split_cols(wcat, 1)
    return torch.bmm(acat.transpose(1, 2), bcat.t().unsqueeze(0).expand(V, -1, -1), out_dtype=torch.float32)


for name, fn in (('tcgen05 fused', tc), ('split + library bf16 GEMM', lib)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    by = 4.0 * V * S * (C + N)
    print(f'{name:28s} {ms * 1e3:8.1f} us   {by / ms / 1e6:8.1f} GB/s algorithmic   {2.0 * V * S * C * N / ms / 1e9:8.1f} TFLOP/s (fp32-equivalent)')

# data gradient and weight gradient of the same level
gvg = torch.randn(V, S, N, device='cuda')
gfeat = torch.zeros(V, C, H0, W0, device='cuda')
wpack_t = SF.pack_weight_tc(wcat.t().contiguous())
gw = torch.empty(N, C, device='cuda')
from sgcdet_b200 import _lib
scratch = torch.empty(_lib.load().sgc_project_tc_wgrad_scratch_floats(N, C), device='cuda')


def bwd_data():
    call('sgc_project_tc_bwd_data', ptr(gvg), V, S, N, ptr(wpack_t), C, ptr(gfeat), H0 * W0, stream())


def wgrad():
    call('sgc_project_tc_wgrad', ptr(gvg), ptr(feat), H0 * W0, V, S, N, C, ptr(gw), ptr(scratch), stream())


for name, fn in (('tcgen05 data gradient', bwd_data), ('tcgen05 weight gradient', wgrad)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    by = 4.0 * V * S * (C + N)
    print(f'{name:28s} {ms * 1e3:8.1f} us   {by / ms / 1e6:8.1f} GB/s algorithmic')
