#!/bin/bash
# Round 2, session AE (1 GPU): validation of the pruned build (losing kernel variants and experiment toggles removed).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/r2ae_times.log; }
ts start
timeout 300 python __graft_entry__.py smoke > $O/r2ae_smoke.log 2>&1
ts smoke "$(tail -1 $O/r2ae_smoke.log)"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/r2ae_suite.log
ts suite "$(tail -1 $O/r2ae_suite.log)"
timeout 900 python bench.py > $O/r2ae_bench_n1.json 2> $O/r2ae_bench_n1.err
ts bench-default "$(python -c "import json;d=json.load(open('$O/r2ae_bench_n1.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'],d['loss_vs_oracle_rel'],d['gpu_launches_per_step'])" 2>&1 | tail -1)"
timeout 300 python bench.py --steps 300 --no-cpu-baseline --no-reference-gpu --no-view-sharded --no-train-step --skip-e2e > $O/r2ae_bench_n1_300.json 2> $O/r2ae_bench_n1_300.err
ts bench-300 "$(python -c "import json;d=json.load(open('$O/r2ae_bench_n1_300.json'));print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"
