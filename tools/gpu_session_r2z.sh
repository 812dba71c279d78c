#!/bin/bash
# Round 2, session Z (1 GPU): suite; four-row gathers in the cross-view kernels of the small levels (A/B), backward issued before
# the loss value (A/B).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/r2z_times.log; }
ts start
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > $O/r2z_suite.log
ts suite "$(tail -1 $O/r2z_suite.log)"
B="timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --no-view-sharded --no-train-step --skip-e2e --steps 300"
run() { name=$1; shift; env "$@" $B > $O/r2z_ab_$name.json 2> $O/r2z_ab_$name.err; ts ab-$name "$(python -c "import json;d=json.load(open('$O/r2z_ab_$name.json'));print(d['value'],d['ms_per_step'],d['gpu_launches_per_step'],d['loss'])" 2>&1 | tail -1)"; }
run def_1 SGC_X=1
run cv0_1 SGC_CV_SMALL_Q=0
run mainloss_1 SGC_LOSS_ON_SIDE=0
run def_2 SGC_X=1
run cv0_2 SGC_CV_SMALL_Q=0
run mainloss_2 SGC_LOSS_ON_SIDE=0
run cvall SGC_CV_SMALL_Q=100000
run def_3 SGC_X=1
SGC_GRAPH_TRACE=$O/r2z_trace.json timeout 300 python tools/profile_step.py > $O/r2z_profile.txt 2>&1
python tools/graph_timeline.py $O/r2z_trace.json 30 $O/r2z_timeline_all.txt > $O/r2z_timeline.txt 2>&1
rm -f $O/r2z_trace.json
ts timeline "$(head -1 $O/r2z_timeline.txt)"
