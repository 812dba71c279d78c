#!/bin/bash
# Round 2, session W (1 GPU): full suite (two-launch pair list, unrolled occupancy loss); confirmation of the session-V defaults;
# CTA-count variants; kernel table; timelines of the ScanNet and the "-L" step.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/r2w_times.log; }
ts start
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > $O/r2w_suite.log
ts suite "$(tail -1 $O/r2w_suite.log)"
B="timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --no-view-sharded --no-train-step --skip-e2e --steps 300"
run() { name=$1; shift; env "$@" $B > $O/r2w_ab_$name.json 2> $O/r2w_ab_$name.err; ts ab-$name "$(python -c "import json;d=json.load(open('$O/r2w_ab_$name.json'));print(d['value'],d['ms_per_step'],d['gpu_launches_per_step'])" 2>&1 | tail -1)"; }
run def_1 SGC_X=1
run small0 SGC_ROWS_SMALL_WORKS=0
run small12 SGC_ROWS_SMALL_WORKS=12
run small24 SGC_ROWS_SMALL_WORKS=24
run def_2 SGC_X=1
run cap140 SGC_TC_MAX_CTAS_FWD=140
run cap148 SGC_TC_MAX_CTAS_FWD=148
run def_3 SGC_X=1
timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --no-view-sharded --no-train-step --steps 100 > $O/r2w_full.json 2> $O/r2w_full.err
ts full "$(python -c "import json;d=json.load(open('$O/r2w_full.json'));print(d['value'],d['ms_per_step'],d['roofline']['frac'],[(k,v['avg_ms']) for k,v in d['kernels'].items()])" 2>&1 | tail -1)"
SGC_GRAPH_TRACE=$O/r2w_trace.json timeout 300 python tools/profile_step.py > $O/r2w_profile.txt 2>&1
python tools/graph_timeline.py $O/r2w_trace.json 30 $O/r2w_timeline_all.txt > $O/r2w_timeline.txt 2>&1
rm -f $O/r2w_trace.json
ts timeline "$(head -1 $O/r2w_timeline.txt)"
SGC_GRAPH_TRACE=$O/r2w_trace_l.json timeout 300 python tools/profile_step.py SGCDet_large_ScanNet200 > $O/r2w_profile_large.txt 2>&1
python tools/graph_timeline.py $O/r2w_trace_l.json 30 $O/r2w_timeline_large_all.txt > $O/r2w_timeline_large.txt 2>&1
rm -f $O/r2w_trace_l.json
ts timeline-large "$(head -1 $O/r2w_timeline_large.txt)"
