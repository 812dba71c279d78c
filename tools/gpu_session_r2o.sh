#!/bin/bash
# Round 2, session O (2 GPUs): rows_gemm timeline without the dual converters, gradient averager with bounded CTAs at N=2.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/r2o_times.log; }
ts start


timeout 200 python -m pytest tests/test_gpu_peer.py tests/test_gpu_path.py -q -k "averager or simulated or valid_pyramid" 2>&1 | tail -30 > $O/r2o_tests.log
ts tests "$(tail -1 $O/r2o_tests.log)"
T="timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
A="--gpus 2 --steps 300 --no-cpu-baseline --no-view-sharded --no-train-step --skip-e2e"
run() { name=$1; port=$2; shift; shift; $T --master-port $port bench.py $A "$@" > $O/r2o_n2_$name.json 2> $O/r2o_n2_$name.err; ts n2-$name "rc=$? $(python -c "import json;d=json.loads(open('$O/r2o_n2_$name.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['gpu_launches_per_step'])" 2>&1 | tail -1)"; }
run peer1 29551
run nccl1 29552 --grad-allreduce nccl
run noar1 29553 --no-grad-allreduce



tail -8 $O/r2o_n2_peer1.err > $O/r2o_n2_peer1_tail.txt
