#!/bin/bash
# Round 2, session AF (8 GPUs): scene-batch DP at N=8 with the final build -- in-graph peer-memory gradient average vs none,
# and the full line (e2e + view-sharded leg).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
N=${1:-8}
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/r2af_times.log; }
ts start N=$N
T="timeout -k 5 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
A="--gpus $N --steps 200 --no-cpu-baseline --no-view-sharded --no-train-step --skip-e2e"
run() { name=$1; port=$2; shift; shift; $T --master-port $port bench.py $A "$@" > $O/r2af_n${N}_$name.json 2> $O/r2af_n${N}_$name.err; ts n$N-$name "rc=$? $(python -c "import json;d=json.loads(open('$O/r2af_n${N}_$name.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['gpu_launches_per_step'])" 2>&1 | tail -1)"; }
run peer 29581
run noar 29583 --no-grad-allreduce
$T --master-port 29584 bench.py --gpus $N --steps 50 --no-cpu-baseline --no-train-step > $O/r2af_n${N}_full.json 2> $O/r2af_n${N}_full.err
ts n$N-full "rc=$? $(python -c "import json;d=json.loads(open('$O/r2af_n${N}_full.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['view_sharded'])" 2>&1 | tail -1)"
tail -8 $O/r2af_n${N}_full.err > $O/r2af_n${N}_full_tail.txt
