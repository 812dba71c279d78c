#!/bin/bash
# Round 2, session C: grouped weight-gradient launch, nibble-histogram top-k; variants benched side by side.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/r2c_times.log; }
ts start
timeout 300 python -m pytest tests/test_gpu_rows_gemm.py tests/test_gpu_path.py -x -q -k "wgrad or grads or topk" 2>&1 | tail -15 > $O/r2c_newkernels.log
ts new-kernels "$(tail -1 $O/r2c_newkernels.log)"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/r2c_suite.log
ts suite "$(tail -1 $O/r2c_suite.log)"
B="timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --skip-e2e --steps 200"
run() { name=$1; shift; env "$@" $B > $O/r2c_bench_$name.json 2> $O/r2c_bench_$name.err; ts bench-$name "$(python -c "import json;d=json.load(open('$O/r2c_bench_$name.json'));print(d['value'],d['ms_per_step'],d['gpu_launches_per_step'])" 2>&1 | tail -1)"; }
run default X=1
run nogroup SGC_WGRAD_GROUP=0
run oldtopk SGC_TOPK_GRID=0
run default2 X=1
timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --skip-e2e --steps 100 --scenes-per-gpu 4 > $O/r2c_bench_b4.json 2> $O/r2c_bench_b4.err
ts bench-b4 "$(python -c "import json;d=json.load(open('$O/r2c_bench_b4.json'));print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"
timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --skip-e2e --steps 50 --config SGCDet_large_ScanNet200 > $O/r2c_bench_large.json 2> $O/r2c_bench_large.err
ts bench-large "$(python -c "import json;d=json.load(open('$O/r2c_bench_large.json'));print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"
SGC_GRAPH_TRACE=$O/r2c_trace.json timeout 300 python tools/profile_step.py > $O/r2c_profile_step.txt 2>&1
python tools/graph_timeline.py $O/r2c_trace.json 30 $O/r2c_timeline_all.txt > $O/r2c_timeline.txt 2>&1
rm -f $O/r2c_trace.json
ts timeline
