#!/bin/bash
# Round 2, session N (1 GPU): sgc_rows_gemm_tc variants A/B in ONE box (box-to-box variation is +-1.5 %).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/r2n_times.log; }
ts start
timeout 300 python -m pytest tests/test_gpu_rows_gemm.py -q -x 2>&1 | tail -30 > $O/r2n_tests_gemm.log
ts tests-gemm "$(tail -1 $O/r2n_tests_gemm.log)"
timeout 200 python tools/rows_gemm_timeline.py 6400 256 256 > $O/r2n_rows_gemm_timeline.txt 2>&1
ts timeline "$(grep -c us $O/r2n_rows_gemm_timeline.txt)"
B="timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --no-view-sharded --no-train-step --skip-e2e --steps 300"
run() { name=$1; shift; env "$@" $B > $O/r2n_ab_$name.json 2> $O/r2n_ab_$name.err; ts ab-$name "$(python -c "import json;d=json.load(open('$O/r2n_ab_$name.json'));print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"; }
run dual1 X=1
run dual0 SGC_ROWS_DUAL=0
run dual1_nb4 SGC_ROWS_NB=4
run dual0_nb4 SGC_ROWS_DUAL=0 SGC_ROWS_NB=4
run dual1_b X=1
run dual0_b SGC_ROWS_DUAL=0
