"""Where the ~10 us of one sgc_rows_gemm_tc launch go: clock64 stamps of CTA 0 (sgc_rows_gemm_tc_set_debug).
   python tools/rows_gemm_timeline.py [R K N]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from sgcdet_b200 import _lib, build, functional as SF  # noqa: E402

build.build()
lib = _lib.load()
NAMES = ['entry', 'barriers initialised', 'TMEM allocated + CTA sync', 'first TMA issued', 'first tile landed',
         'first k-slab converted', 'last k-slab converted', "first slab's MMAs issued", 'accumulator committed',
         'epilogue sees accumulator', 'last store issued', 'stores drained', 'TMEM freed']
shapes = [tuple(int(v) for v in sys.argv[1:4])] if len(sys.argv) >= 4 else [(6400, 256, 256), (400, 256, 256), (6400, 256, 512),
                                                                              (6400, 512, 256)]
dev = torch.device('cuda')
stamps = torch.zeros(16, device=dev, dtype=torch.int64)
mhz = torch.cuda.clock_rate() / 1e3 if hasattr(torch.cuda, 'clock_rate') else 1965.0
for R, K, N in shapes:
    x = torch.randn(R, K, device=dev)
    w = torch.randn(N, K, device=dev)
    pk = SF.pack_weight_tc(w)
    b = torch.randn(N, device=dev)
    for _ in range(3):
        SF.rows_linear(x, pk, N, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        SF.rows_linear(x, pk, N, b)
    e1.record()
    torch.cuda.synchronize()
    lib.sgc_rows_gemm_tc_set_debug(stamps.data_ptr())
    SF.rows_linear(x, pk, N, b)
    torch.cuda.synchronize()
    lib.sgc_rows_gemm_tc_set_debug(None)
    t = stamps.cpu().tolist()
    n_cta = lib.sgc_rows_gemm_tc_auto_ncta(R, N, 1)
    print(f'R={R} K={K} N={N}: n_cta={n_cta}, {e0.elapsed_time(e1) / 20 * 1e3:.2f} us per back-to-back launch (events)')
    for i, name in enumerate(NAMES):
        print(f'   {(t[i] - t[0]) / mhz:8.2f} us  {name}')
