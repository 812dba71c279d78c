#!/bin/bash
# Round 2, session I (2 GPUs): peer-memory collectives between two real GPUs -- scene-batch DP with the gradient average
# inside the CUDA graph vs NCCL after the replay, the view-sharded leg (config 5) and the train-step leg (config 3) at N=2;
# plus the single-GPU peer tests (simulated ranks, two processes on one GPU) and two scheduling A/B runs.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/r2i_times.log; }
ts start
timeout 600 python -m pytest tests/test_gpu_peer.py -q 2>&1 | tail -40 > $O/r2i_peer_tests.log
ts peer-tests "$(tail -1 $O/r2i_peer_tests.log)"
T="timeout -k 5 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$T --master-port 29521 bench.py --gpus 2 --steps 100 --no-cpu-baseline > $O/r2i_n2_peer.json 2> $O/r2i_n2_peer.err
ts n2-peer "rc=$? $(python -c "import json;d=json.loads(open('$O/r2i_n2_peer.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e'],d['config']['parallelism'],d['view_sharded'],d['train_step'])" 2>&1 | tail -1)"
$T --master-port 29522 bench.py --gpus 2 --steps 100 --no-cpu-baseline --no-view-sharded --no-train-step --skip-e2e --grad-allreduce nccl > $O/r2i_n2_nccl.json 2> $O/r2i_n2_nccl.err
ts n2-nccl "rc=$? $(python -c "import json;d=json.loads(open('$O/r2i_n2_nccl.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"
$T --master-port 29523 bench.py --gpus 2 --steps 100 --no-cpu-baseline --no-view-sharded --no-train-step --skip-e2e --no-grad-allreduce > $O/r2i_n2_noar.json 2> $O/r2i_n2_noar.err
ts n2-no-allreduce "rc=$? $(python -c "import json;d=json.loads(open('$O/r2i_n2_noar.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"
$T --master-port 29524 bench.py --gpus 2 --steps 10 --no-cpu-baseline --no-train-step --skip-e2e --view-sharded-views 100 > $O/r2i_n2_vs100.json 2> $O/r2i_n2_vs100.err
ts n2-view-sharded-v100 "rc=$? $(python -c "import json;d=json.loads(open('$O/r2i_n2_vs100.json').read().strip().splitlines()[-1]);print(d['view_sharded'])" 2>&1 | tail -1)"
B="timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --no-view-sharded --no-train-step --skip-e2e --steps 200"
run() { name=$1; shift; env "$@" $B > $O/r2i_ab_$name.json 2> $O/r2i_ab_$name.err; ts ab-$name "$(python -c "import json;d=json.load(open('$O/r2i_ab_$name.json'));print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"; }
run base X=1
run fwdcap96 SGC_TC_MAX_CTAS_FWD=96
run fwdcap64 SGC_TC_MAX_CTAS_FWD=64
run minb3 SGC_LIFT_MINB=3
for f in $O/r2i_n2_peer.err $O/r2i_n2_vs100.err; do tail -8 $f > ${f%.err}_tail.txt; done
