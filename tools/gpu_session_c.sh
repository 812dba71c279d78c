#!/bin/bash
# GPU session C: tensor-core weight-gradient kernel, SM caps.  Outputs in gpurun_out/.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/c_times.log; }
ts start
timeout 300 python -m pytest tests/test_gpu_rows_gemm.py tests/test_gpu_rowops.py -q 2>&1 | tail -40 > $O/c_rows.log
ts rows-tests "$(tail -1 $O/c_rows.log)"
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_rows_gemm.py --deselect tests/test_gpu_rowops.py 2>&1 | tail -30 > $O/c_tests.log
ts full-tests "$(tail -1 $O/c_tests.log)"
B="timeout 300 python bench.py --no-cpu-baseline --skip-e2e"
run() { name=$1; shift; env "$@" $B > $O/c_bench_$name.json 2> $O/c_bench_$name.err; ts bench-$name "$(python -c "import json;d=json.load(open('$O/c_bench_$name.json'));print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"; }
run default X=1
run wgrad0 SGC_ROWS_WGRAD_TC=0
run fwd132 SGC_TC_MAX_CTAS_FWD=132
run fwd116_all140 SGC_TC_MAX_CTAS_FWD=116 SGC_TC_MAX_CTAS=140
run fwd132_all140 SGC_TC_MAX_CTAS_FWD=132 SGC_TC_MAX_CTAS=140
run all140 SGC_TC_MAX_CTAS=140
run all132 SGC_TC_MAX_CTAS=132
env X=1 $B --scenes-per-gpu 4 > $O/c_bench_b4.json 2> $O/c_bench_b4.err; ts bench-b4 "$(python -c "import json;d=json.load(open('$O/c_bench_b4.json'));print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"
SGC_GRAPH_TRACE=$O/c_trace.json timeout 300 python tools/profile_step.py > $O/c_profile_step.txt 2>&1
python tools/graph_timeline.py $O/c_trace.json 20 $O/c_timeline_all.txt > $O/c_timeline.txt 2>&1
rm -f $O/c_trace.json
ts timeline
