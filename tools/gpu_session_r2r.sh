#!/bin/bash
# Round 2, session R (1 GPU): full suite; A/B of the weight-gradient stream priority in one box; timelines.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/r2r_times.log; }
ts start
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/r2r_suite.log
ts suite "$(tail -1 $O/r2r_suite.log)"
B="timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --no-view-sharded --no-train-step --skip-e2e --steps 300"
run() { name=$1; shift; env "$@" $B > $O/r2r_ab_$name.json 2> $O/r2r_ab_$name.err; ts ab-$name "$(python -c "import json;d=json.load(open('$O/r2r_ab_$name.json'));print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"; }
run prio_hi_1 SGC_WSTREAM_PRIO=-1
run prio_def_1 SGC_WSTREAM_PRIO=0
run prio_hi_2 SGC_WSTREAM_PRIO=-1
run prio_def_2 SGC_WSTREAM_PRIO=0
run prio_hi_3 SGC_WSTREAM_PRIO=-1
run prio_def_3 SGC_WSTREAM_PRIO=0
SGC_GRAPH_TRACE=$O/r2r_trace.json timeout 300 python tools/profile_step.py > $O/r2r_profile.txt 2>&1
python tools/graph_timeline.py $O/r2r_trace.json 30 $O/r2r_timeline_all.txt > $O/r2r_timeline.txt 2>&1
rm -f $O/r2r_trace.json
ts timeline "$(head -1 $O/r2r_timeline.txt)"
