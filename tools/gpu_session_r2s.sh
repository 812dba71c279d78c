#!/bin/bash
# Round 2, session S (2 GPUs): gradient averager with the one-launch tail at N=2 (vs NCCL, vs none), view-sharded + train legs.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/r2s_times.log; }
ts start
timeout 200 python -m pytest tests/test_gpu_peer.py tests/test_gpu_project_tc.py -q 2>&1 | tail -20 > $O/r2s_tests.log
ts tests "$(tail -1 $O/r2s_tests.log)"
T="timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
A="--gpus 2 --steps 300 --no-cpu-baseline --no-view-sharded --no-train-step --skip-e2e"
run() { name=$1; port=$2; shift; shift; $T --master-port $port bench.py $A "$@" > $O/r2s_n2_$name.json 2> $O/r2s_n2_$name.err; ts n2-$name "rc=$? $(python -c "import json;d=json.loads(open('$O/r2s_n2_$name.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['gpu_launches_per_step'])" 2>&1 | tail -1)"; }
run peer1 29571
run nccl1 29572 --grad-allreduce nccl
run noar1 29573 --no-grad-allreduce
run peer2 29574
$T --master-port 29575 bench.py --gpus 2 --steps 100 --no-cpu-baseline > $O/r2s_n2_full.json 2> $O/r2s_n2_full.err
ts n2-full "rc=$? $(python -c "import json;d=json.loads(open('$O/r2s_n2_full.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['view_sharded'],d['train_step'])" 2>&1 | tail -1)"
tail -8 $O/r2s_n2_full.err > $O/r2s_n2_full_tail.txt
