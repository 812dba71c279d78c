#!/bin/bash
# Round 2, session N (1 GPU): sgc_rows_gemm_tc with dual converter groups / per-warp stores / early producers: parity, timeline, bench.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/r2n_times.log; }
ts start
timeout 300 python -m pytest tests/test_gpu_rows_gemm.py -q -x 2>&1 | tail -30 > $O/r2n_tests_gemm.log
ts tests-gemm "$(tail -1 $O/r2n_tests_gemm.log)"
timeout 200 python tools/rows_gemm_timeline.py > $O/r2n_rows_gemm_timeline.txt 2>&1
ts timeline "$(grep -c us $O/r2n_rows_gemm_timeline.txt)"
timeout 600 python -m pytest tests/test_gpu_path.py tests/test_gpu_full_shape_parity.py tests/test_gpu_peer.py -q 2>&1 | tail -30 > $O/r2n_tests_path.log
ts tests-path "$(tail -1 $O/r2n_tests_path.log)"
B="timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --no-view-sharded --no-train-step --skip-e2e --steps 200"
for i in 1 2; do
$B > $O/r2n_bench_$i.json 2> $O/r2n_bench_$i.err
ts bench-$i "$(python -c "import json;d=json.load(open('$O/r2n_bench_$i.json'));print(d['value'],d['ms_per_step'],d['gpu_launches_per_step'])" 2>&1 | tail -1)"
done
$B --config SGCDet_large_ScanNet200 --steps 50 > $O/r2n_bench_large.json 2> $O/r2n_bench_large.err
ts bench-large "$(python -c "import json;d=json.load(open('$O/r2n_bench_large.json'));print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"
