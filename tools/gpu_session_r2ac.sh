#!/bin/bash
# Round 2, session AC (2 GPUs): scene-batch DP with the final step (own dropout masks, loss on the loss stream): weight-stream
# priority A/B under the gradient averager, the full N=2 line with the view-sharded and train-step legs.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/r2ac_times.log; }
ts start
timeout 200 python -m pytest tests/test_gpu_peer.py -q 2>&1 | tail -20 > $O/r2ac_tests.log
ts tests "$(tail -1 $O/r2ac_tests.log)"
T="timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
A="--gpus 2 --steps 300 --no-cpu-baseline --no-view-sharded --no-train-step --skip-e2e"
run() { name=$1; port=$2; shift; shift; env $ENVV $T --master-port $port bench.py $A "$@" > $O/r2ac_n2_$name.json 2> $O/r2ac_n2_$name.err; ts n2-$name "rc=$? $(python -c "import json;d=json.loads(open('$O/r2ac_n2_$name.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['gpu_launches_per_step'])" 2>&1 | tail -1)"; }
ENVV="SGC_X=1" run peer1 29571
ENVV="SGC_WSTREAM_PRIO=0" run peer_wprio0_1 29572
ENVV="SGC_X=1" run peer2 29573
ENVV="SGC_WSTREAM_PRIO=0" run peer_wprio0_2 29574
ENVV="SGC_X=1" run noar 29575 --no-grad-allreduce
ENVV="SGC_X=1" $T --master-port 29576 bench.py --gpus 2 --steps 100 --no-cpu-baseline > $O/r2ac_n2_full.json 2> $O/r2ac_n2_full.err
ts n2-full "rc=$? $(python -c "import json;d=json.loads(open('$O/r2ac_n2_full.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['view_sharded'],d['train_step'])" 2>&1 | tail -1)"
tail -8 $O/r2ac_n2_full.err > $O/r2ac_n2_full_tail.txt
