#!/bin/bash
# GPU session D: loss value beside the backward, aggregated top-k histogram, forward cap default.  Outputs in gpurun_out/.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/d_times.log; }
ts start
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > $O/d_tests.log
ts full-tests "$(tail -1 $O/d_tests.log)"
B="timeout 300 python bench.py --no-cpu-baseline --skip-e2e --steps 100"
run() { name=$1; shift; env "$@" $B > $O/d_bench_$name.json 2> $O/d_bench_$name.err; ts bench-$name "$(python -c "import json;d=json.load(open('$O/d_bench_$name.json'));print(d['value'],d['ms_per_step'],d['loss'])" 2>&1 | tail -1)"; }
run default X=1
run default2 X=1
run fwd0 SGC_TC_MAX_CTAS_FWD=0
run fwd116 SGC_TC_MAX_CTAS_FWD=116
run all132 SGC_TC_MAX_CTAS=132
run lvl0 SGC_LEVEL_STREAMS=0
run noprezero SGC_PREZERO=0
timeout 300 python bench.py --no-cpu-baseline > $O/d_bench_full.json 2> $O/d_bench_full.err
ts bench-full "$(python -c "import json;d=json.load(open('$O/d_bench_full.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['kernel'])" 2>&1 | tail -1)"
SGC_GRAPH_TRACE=$O/d_trace.json timeout 300 python tools/profile_step.py > $O/d_profile_step.txt 2>&1
python tools/graph_timeline.py $O/d_trace.json 20 $O/d_timeline_all.txt > $O/d_timeline.txt 2>&1
rm -f $O/d_trace.json
ts timeline
