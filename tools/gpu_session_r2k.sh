#!/bin/bash
# Round 2, session K (4 GPUs): scene-batch DP with the in-graph two-shot gradient average, view-sharded leg at N=4.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
N=${1:-4}
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/r2k_times.log; }
ts start N=$N
T="timeout -k 5 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$T --master-port 29531 bench.py --gpus $N --steps 100 --no-cpu-baseline --no-train-step > $O/r2k_n${N}_peer.json 2> $O/r2k_n${N}_peer.err
ts n$N-peer "rc=$? $(python -c "import json;d=json.loads(open('$O/r2k_n${N}_peer.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e'],d['view_sharded'])" 2>&1 | tail -1)"
$T --master-port 29532 bench.py --gpus $N --steps 100 --no-cpu-baseline --no-view-sharded --no-train-step --skip-e2e --grad-allreduce nccl > $O/r2k_n${N}_nccl.json 2> $O/r2k_n${N}_nccl.err
ts n$N-nccl "rc=$? $(python -c "import json;d=json.loads(open('$O/r2k_n${N}_nccl.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"
$T --master-port 29533 bench.py --gpus $N --steps 100 --no-cpu-baseline --no-view-sharded --no-train-step --skip-e2e --no-grad-allreduce > $O/r2k_n${N}_noar.json 2> $O/r2k_n${N}_noar.err
ts n$N-no-allreduce "rc=$? $(python -c "import json;d=json.loads(open('$O/r2k_n${N}_noar.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"
tail -8 $O/r2k_n${N}_peer.err > $O/r2k_n${N}_peer_tail.txt
