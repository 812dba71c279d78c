#!/bin/bash
# Round 2, session F: bias reduce folded into lift_bwd, lane-parallel DFA3D operator kernels, harness, parity outliers.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/r2f_times.log; }
ts start
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/r2f_suite.log
ts suite "$(tail -1 $O/r2f_suite.log)"
B="timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --no-view-sharded --no-train-step --skip-e2e --steps 200"
run() { name=$1; shift; env "$@" $B > $O/r2f_bench_$name.json 2> $O/r2f_bench_$name.err; ts bench-$name "$(python -c "import json;d=json.load(open('$O/r2f_bench_$name.json'));print(d['value'],d['ms_per_step'],d['gpu_launches_per_step'])" 2>&1 | tail -1)"; }
run default X=1
run prio SGC_CHAIN_PRIO=0,0,-1
run default2 X=1
timeout 600 python bench.py --steps 20 > $O/r2f_bench_full.json 2> $O/r2f_bench_full.err
ts bench-full "$(python -c "import json;d=json.load(open('$O/r2f_bench_full.json'));print(d['value'],d['e2e']['value'],d['operator_bench'],d['train_step'],d['view_sharded'],d['reference_gpu'])" 2>&1 | tail -1)"
SGC_GRAPH_TRACE=$O/r2f_trace.json timeout 300 python tools/profile_step.py > $O/r2f_profile_step.txt 2>&1
python tools/graph_timeline.py $O/r2f_trace.json 30 $O/r2f_timeline_all.txt > $O/r2f_timeline.txt 2>&1
rm -f $O/r2f_trace.json
ts timeline
