#!/bin/bash
# GPU session K: hybrid top-k (bitwise leading bits + histogram passes): parity, duration, bench.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/k_times.log; }
ts start
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > $O/k_tests.log
ts full-tests "$(tail -1 $O/k_tests.log)"
python - > $O/k_topk_times.txt 2>&1 <<'PY'
import torch, sys
sys.path.insert(0, '.')
from sgcdet_b200 import functional as SF
for N, k in ((3200, 800), (25600, 6400)):
    occ = torch.sigmoid(torch.randn(N, device='cuda'))
    for _ in range(5):
        SF.topk_select(occ, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        SF.topk_select(occ, k)
    e1.record()
    torch.cuda.synchronize()
    print(N, k, 'us per call (incl. launch):', e0.elapsed_time(e1) / 200 * 1e3)
PY
ts topk "$(tr '\n' ' ' < $O/k_topk_times.txt)"
B="timeout 300 python bench.py --no-cpu-baseline --skip-e2e --steps 200"
for rep in 1 2 3; do
$B > $O/k_bench_default_$rep.json 2> $O/k_bench_default_$rep.err
ts bench-default_$rep "$(python -c "import json;d=json.load(open('$O/k_bench_default_$rep.json'));print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"
done
