#!/bin/bash
# Round 2, session L (1 GPU): the full GPU suite after the legacy-path prune + the bucketed gradient averager, bench line.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/r2l_times.log; }
ts start
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/r2l_suite.log
ts suite "$(tail -1 $O/r2l_suite.log)"
timeout 300 python __graft_entry__.py smoke > $O/r2l_smoke.log 2>&1
ts smoke "$(tail -1 $O/r2l_smoke.log)"
timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --no-view-sharded --no-train-step --skip-e2e --steps 200 > $O/r2l_bench_n1.json 2> $O/r2l_bench_n1.err
ts bench "$(python -c "import json;d=json.load(open('$O/r2l_bench_n1.json'));print(d['value'],d['ms_per_step'],d['gpu_launches_per_step'])" 2>&1 | tail -1)"
timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --no-view-sharded --no-train-step --skip-e2e --steps 50 --config SGCDet_large_ScanNet200 > $O/r2l_bench_large.json 2> $O/r2l_bench_large.err
ts bench-large "$(python -c "import json;d=json.load(open('$O/r2l_bench_large.json'));print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"
tail -5 $O/r2l_bench_n1.err > $O/r2l_bench_n1_tail.txt
