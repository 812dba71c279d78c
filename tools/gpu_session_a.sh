#!/bin/bash
# One GPU session: new rows-GEMM kernel tests, full GPU suite, bench variants, graph timeline.  Outputs in gpurun_out/.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/a_times.log; }
ts start
timeout 300 python -m pytest tests/test_gpu_rows_gemm.py -q 2>&1 | tail -40 > $O/a_rows.log
ts rows-tests rc=$?
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_rows_gemm.py 2>&1 | tail -40 > $O/a_tests.log
ts full-tests
if ! grep -q " passed" $O/a_tests.log || grep -q "failed" $O/a_tests.log; then
  SGC_ROWS_TC=0 timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_rows_gemm.py 2>&1 | tail -40 > $O/a_tests_rows0.log
  ts full-tests-rows0
fi
timeout 400 python bench.py --no-cpu-baseline > $O/a_bench_default.json 2> $O/a_bench_default.err
ts bench-default
SGC_ROWS_TC=0 timeout 300 python bench.py --no-cpu-baseline --skip-e2e > $O/a_bench_rows0.json 2> $O/a_bench_rows0.err
ts bench-rows0
SGC_TC_MAX_CTAS=132 timeout 300 python bench.py --no-cpu-baseline --skip-e2e > $O/a_bench_cap132.json 2> $O/a_bench_cap132.err
ts bench-cap132
SGC_TC_MAX_CTAS=140 timeout 300 python bench.py --no-cpu-baseline --skip-e2e > $O/a_bench_cap140.json 2> $O/a_bench_cap140.err
ts bench-cap140
SGC_GRAPH_TRACE=$O/a_trace.json timeout 300 python tools/profile_step.py > $O/a_profile_step.txt 2>&1
python tools/graph_timeline.py $O/a_trace.json 20 $O/a_timeline_all.txt > $O/a_timeline.txt 2>&1
rm -f $O/a_trace.json
ts timeline
