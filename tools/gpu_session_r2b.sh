#!/bin/bash
# Round 2, session B: grid top-k + occupancy-loss kernels, forward prepare chaining; variants benched side by side.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/r2b_times.log; }
ts start
timeout 180 python -m pytest tests/test_gpu_path.py -x -q -k "topk or occ_loss" 2>&1 | tail -15 > $O/r2b_newkernels.log
ts new-kernels "$(tail -1 $O/r2b_newkernels.log)"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/r2b_suite.log
ts suite "$(tail -1 $O/r2b_suite.log)"
B="timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --skip-e2e --steps 200"
run() { name=$1; shift; env "$@" $B > $O/r2b_bench_$name.json 2> $O/r2b_bench_$name.err; ts bench-$name "$(python -c "import json;d=json.load(open('$O/r2b_bench_$name.json'));print(d['value'],d['ms_per_step'],d['gpu_launches_per_step'])" 2>&1 | tail -1)"; }
run default X=1
run nochain SGC_CHAIN_PREPARE=0
run oldtopk SGC_TOPK_GRID=0
run fwd116 SGC_TC_MAX_CTAS_FWD=116
run fwd100 SGC_TC_MAX_CTAS_FWD=100
run fwd148 SGC_TC_MAX_CTAS_FWD=148
run conn1 CUDA_DEVICE_MAX_CONNECTIONS=1
run conn1_fwd116 CUDA_DEVICE_MAX_CONNECTIONS=1 SGC_TC_MAX_CTAS_FWD=116
run default2 X=1
SGC_GRAPH_TRACE=$O/r2b_trace.json timeout 300 python tools/profile_step.py > $O/r2b_profile_step.txt 2>&1
python tools/graph_timeline.py $O/r2b_trace.json 30 $O/r2b_timeline_all.txt > $O/r2b_timeline.txt 2>&1
rm -f $O/r2b_trace.json
ts timeline
