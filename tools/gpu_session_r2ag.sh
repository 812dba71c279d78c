#!/bin/bash
# Round 2, session AG (1 GPU): last A/B of the round -- column width of the grouped weight-gradient kernel, resident CTAs of the
# C = 128 lift backward ("-L" configs).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/r2ag_times.log; }
ts start
timeout 200 python -m pytest tests/test_gpu_rows_gemm.py -q -x 2>&1 | tail -3 > $O/r2ag_tests_def.log
SGC_RW_NCTA=256 timeout 200 python -m pytest tests/test_gpu_rows_gemm.py -q -x 2>&1 | tail -3 > $O/r2ag_tests_rw256.log
ts tests "$(tail -1 $O/r2ag_tests_def.log) / $(tail -1 $O/r2ag_tests_rw256.log)"
B="timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --no-view-sharded --no-train-step --skip-e2e --steps 300"
run() { name=$1; shift; env "$@" $B > $O/r2ag_ab_$name.json 2> $O/r2ag_ab_$name.err; ts ab-$name "$(python -c "import json;d=json.load(open('$O/r2ag_ab_$name.json'));print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"; }
run def_1 SGC_X=1
run rw256_1 SGC_RW_NCTA=256
run def_2 SGC_X=1
run rw256_2 SGC_RW_NCTA=256
L="timeout 300 python bench.py --config SGCDet_large_ScanNet200 --no-cpu-baseline --no-reference-gpu --no-view-sharded --no-train-step --skip-e2e --steps 100"
runl() { name=$1; shift; env "$@" $L > $O/r2ag_large_$name.json 2> $O/r2ag_large_$name.err; ts large-$name "$(python -c "import json;d=json.load(open('$O/r2ag_large_$name.json'));print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"; }
runl def_1 SGC_X=1
runl minb2 SGC_LIFT_MINB128=2
runl minb4 SGC_LIFT_MINB128=4
runl rw256 SGC_RW_NCTA=256
runl def_2 SGC_X=1
