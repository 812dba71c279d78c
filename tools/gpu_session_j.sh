#!/bin/bash
# GPU session J: tiny256 parity tests, bitwise-search top-k, stdout contract, timeline of the large config.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/j_times.log; }
ts start
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > $O/j_tests.log
ts full-tests "$(tail -1 $O/j_tests.log)"
B="timeout 300 python bench.py --no-cpu-baseline --skip-e2e --steps 200"
for rep in 1 2; do
$B > $O/j_bench_default_$rep.json 2> $O/j_bench_default_$rep.err
ts bench-default_$rep "$(wc -l < $O/j_bench_default_$rep.json) line(s): $(python -c "import json;d=json.load(open('$O/j_bench_default_$rep.json'));print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"
done
SGC_GRAPH_TRACE=$O/j_trace.json timeout 300 python tools/profile_step.py > $O/j_profile_step.txt 2>&1
python tools/graph_timeline.py $O/j_trace.json 30 $O/j_timeline_all.txt > $O/j_timeline.txt 2>&1
SGC_GRAPH_TRACE=$O/j_trace.json timeout 300 python tools/profile_step.py SGCDet_large_ScanNet200 40 > $O/j_profile_step_large.txt 2>&1
python tools/graph_timeline.py $O/j_trace.json 40 $O/j_timeline_large_all.txt > $O/j_timeline_large.txt 2>&1
rm -f $O/j_trace.json
ts timelines
