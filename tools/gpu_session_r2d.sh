#!/bin/bash
# Round 2, session D: tile-binned lift backward, fold / unfold kernels, deeper staging of the grouped weight gradients.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/r2d_times.log; }
ts start
timeout 300 python -m pytest tests/test_gpu_lift_tiles.py tests/test_gpu_rowops.py tests/test_gpu_rows_gemm.py -x -q -k "tile or fold or wgrad or grads" 2>&1 | tail -25 > $O/r2d_newkernels.log
ts new-kernels "$(tail -1 $O/r2d_newkernels.log)"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > $O/r2d_suite.log
ts suite "$(tail -1 $O/r2d_suite.log)"
B="timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --skip-e2e --steps 200"
run() { name=$1; shift; env "$@" $B > $O/r2d_bench_$name.json 2> $O/r2d_bench_$name.err; ts bench-$name "$(python -c "import json;d=json.load(open('$O/r2d_bench_$name.json'));print(d['value'],d['ms_per_step'],d['gpu_launches_per_step'])" 2>&1 | tail -1)"; }
run default X=1
run scatter SGC_LIFT_TILES=0
run default2 X=1
timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --steps 30 > $O/r2d_bench_full.json 2> $O/r2d_bench_full.err
ts bench-full "$(python -c "import json;d=json.load(open('$O/r2d_bench_full.json'));print(d['value'],d['e2e']['value'],d['roofline'],json.dumps(d['kernels']))" 2>&1 | tail -1)"
timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --skip-e2e --steps 50 --config SGCDet_large_ScanNet200 > $O/r2d_bench_large.json 2> $O/r2d_bench_large.err
ts bench-large "$(python -c "import json;d=json.load(open('$O/r2d_bench_large.json'));print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"
SGC_GRAPH_TRACE=$O/r2d_trace.json timeout 300 python tools/profile_step.py > $O/r2d_profile_step.txt 2>&1
python tools/graph_timeline.py $O/r2d_trace.json 30 $O/r2d_timeline_all.txt > $O/r2d_timeline.txt 2>&1
rm -f $O/r2d_trace.json
ts timeline
