#!/bin/bash
# Round 2, session P (1 GPU): graph timeline of the step with and without the gradient averager (single rank).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/r2p_times.log; }
ts start
SGC_GRAPH_TRACE=$O/r2p_trace_base.json timeout 300 python tools/profile_step.py > $O/r2p_profile_base.txt 2>&1
python tools/graph_timeline.py $O/r2p_trace_base.json 30 $O/r2p_timeline_base_all.txt > $O/r2p_timeline_base.txt 2>&1
rm -f $O/r2p_trace_base.json
ts base "$(head -1 $O/r2p_timeline_base.txt)"
SGC_PROFILE_AVERAGER=1 SGC_GRAPH_TRACE=$O/r2p_trace_avg.json timeout 300 python tools/profile_step.py > $O/r2p_profile_avg.txt 2>&1
python tools/graph_timeline.py $O/r2p_trace_avg.json 30 $O/r2p_timeline_avg_all.txt > $O/r2p_timeline_avg.txt 2>&1
rm -f $O/r2p_trace_avg.json
ts avg "$(head -1 $O/r2p_timeline_avg.txt)"
