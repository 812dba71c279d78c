#!/bin/bash
# Round 2, session Q (2 GPUs): gradient averager (high-priority comm stream, small-footprint launches) at N=2, and the timeline
# of the step with the averager on one rank.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/r2q_times.log; }
ts start
T="timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
A="--gpus 2 --steps 300 --no-cpu-baseline --no-view-sharded --no-train-step --skip-e2e"
run() { name=$1; port=$2; shift; shift; $T --master-port $port bench.py $A "$@" > $O/r2q_n2_$name.json 2> $O/r2q_n2_$name.err; ts n2-$name "rc=$? $(python -c "import json;d=json.loads(open('$O/r2q_n2_$name.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['gpu_launches_per_step'])" 2>&1 | tail -1)"; }
run peer1 29561
run nccl1 29562 --grad-allreduce nccl
run noar1 29563 --no-grad-allreduce
SGC_PROFILE_AVERAGER=1 SGC_GRAPH_TRACE=$O/r2q_trace_avg.json timeout 300 python tools/profile_step.py > $O/r2q_profile_avg.txt 2>&1
python tools/graph_timeline.py $O/r2q_trace_avg.json 30 $O/r2q_timeline_avg_all.txt > $O/r2q_timeline_avg.txt 2>&1
rm -f $O/r2q_trace_avg.json
ts avg-timeline "$(head -1 $O/r2q_timeline_avg.txt)"
tail -8 $O/r2q_n2_peer1.err > $O/r2q_n2_peer1_tail.txt
