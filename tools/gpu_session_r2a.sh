#!/bin/bash
# Round 2, session A: full-shape parity tests, bench line with the new legs, CUDA_DEVICE_MAX_CONNECTIONS experiment, timeline.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/r2a_times.log; }
ts start
timeout 600 python -m pytest tests/test_gpu_full_shape_parity.py -x -q 2>&1 | tail -15 > $O/r2a_parity.log
ts parity "$(tail -1 $O/r2a_parity.log)"
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_full_shape_parity.py 2>&1 | tail -5 > $O/r2a_suite.log
ts suite "$(tail -1 $O/r2a_suite.log)"
timeout 500 python bench.py --steps 50 > $O/r2a_bench_n1.json 2> $O/r2a_bench_n1.err
ts bench "$(python -c "import json;d=json.load(open('$O/r2a_bench_n1.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['loss_vs_oracle_rel'],d['reference_gpu'],d['operator_bench'])" 2>&1 | tail -1)"
B="timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --skip-e2e --steps 200"
run() { name=$1; shift; env "$@" $B > $O/r2a_bench_$name.json 2> $O/r2a_bench_$name.err; ts bench-$name "$(python -c "import json;d=json.load(open('$O/r2a_bench_$name.json'));print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"; }
run conn8 CUDA_DEVICE_MAX_CONNECTIONS=8
run conn32 CUDA_DEVICE_MAX_CONNECTIONS=32
run conn1 CUDA_DEVICE_MAX_CONNECTIONS=1

timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --skip-e2e --steps 100 --scenes-per-gpu 4 > $O/r2a_bench_b4.json 2> $O/r2a_bench_b4.err
ts bench-b4 "$(python -c "import json;d=json.load(open('$O/r2a_bench_b4.json'));print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"
for c in 8 32; do
CUDA_DEVICE_MAX_CONNECTIONS=$c SGC_GRAPH_TRACE=$O/r2a_trace.json timeout 300 python tools/profile_step.py > $O/r2a_profile_step_$c.txt 2>&1
python tools/graph_timeline.py $O/r2a_trace.json 30 $O/r2a_timeline_all_conn$c.txt > $O/r2a_timeline_conn$c.txt 2>&1
rm -f $O/r2a_trace.json
ts timeline-$c
done
