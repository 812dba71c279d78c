#!/bin/bash
# Round 2, session E: narrow-head tensor-core products, view-sharded fixes + bench leg, pinned-projection parity.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/r2e_times.log; }
ts start
timeout 300 python -m pytest tests/test_gpu_rows_gemm.py tests/test_gpu_view_sharded.py -x -q -k "narrow or sharded" 2>&1 | tail -25 > $O/r2e_newkernels.log
ts new-kernels "$(tail -1 $O/r2e_newkernels.log)"
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/r2e_suite.log
ts suite "$(tail -1 $O/r2e_suite.log)"
B="timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --no-view-sharded --skip-e2e --steps 200"
run() { name=$1; shift; env "$@" $B > $O/r2e_bench_$name.json 2> $O/r2e_bench_$name.err; ts bench-$name "$(python -c "import json;d=json.load(open('$O/r2e_bench_$name.json'));print(d['value'],d['ms_per_step'],d['gpu_launches_per_step'])" 2>&1 | tail -1)"; }
run default X=1
timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --no-view-sharded --skip-e2e --steps 50 --config SGCDet_large_ScanNet200 > $O/r2e_bench_large.json 2> $O/r2e_bench_large.err
ts bench-large "$(python -c "import json;d=json.load(open('$O/r2e_bench_large.json'));print(d['value'],d['ms_per_step'],d['gpu_launches_per_step'])" 2>&1 | tail -1)"
SGC_HEADS_EXP=0 timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --no-view-sharded --skip-e2e --steps 50 --config SGCDet_large_ScanNet200 > $O/r2e_bench_large_lib.json 2> $O/r2e_bench_large_lib.err
ts bench-large-libheads "$(python -c "import json;d=json.load(open('$O/r2e_bench_large_lib.json'));print(d['value'],d['ms_per_step'],d['gpu_launches_per_step'])" 2>&1 | tail -1)"
timeout 400 python bench.py --no-cpu-baseline --no-reference-gpu --steps 20 > $O/r2e_bench_full.json 2> $O/r2e_bench_full.err
ts bench-full "$(python -c "import json;d=json.load(open('$O/r2e_bench_full.json'));print(d['value'],d['e2e']['value'],d['view_sharded'])" 2>&1 | tail -1)"
SGC_GRAPH_TRACE=$O/r2e_trace.json timeout 300 python tools/profile_step.py > $O/r2e_profile_step.txt 2>&1
python tools/graph_timeline.py $O/r2e_trace.json 30 $O/r2e_timeline_all.txt > $O/r2e_timeline.txt 2>&1
rm -f $O/r2e_trace.json
ts timeline
SGC_GRAPH_TRACE=$O/r2e_trace.json timeout 300 python tools/profile_step.py SGCDet_large_ScanNet200 40 > $O/r2e_profile_step_large.txt 2>&1
python tools/graph_timeline.py $O/r2e_trace.json 30 $O/r2e_timeline_large_all.txt > $O/r2e_timeline_large.txt 2>&1
rm -f $O/r2e_trace.json
ts timeline-large
