#!/bin/bash
# Round 2, session M (2 GPUs): the gradient average issued from the OnStream backward nodes (overlapped, in the graph) vs NCCL
# after the replay vs none; the tests that failed in session L.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/r2m_times.log; }
ts start
timeout 600 python -m pytest tests/test_gpu_peer.py -q -k "averager" 2>&1 | tail -60 > $O/r2m_tests.log
ts tests "$(tail -1 $O/r2m_tests.log)"
T="timeout -k 5 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for i in 1 2; do
$T --master-port 2954$i bench.py --gpus 2 --steps 200 --no-cpu-baseline --no-view-sharded --no-train-step --skip-e2e > $O/r2m_n2_peer$i.json 2> $O/r2m_n2_peer$i.err
ts n2-peer$i "rc=$? $(python -c "import json;d=json.loads(open('$O/r2m_n2_peer$i.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['gpu_launches_per_step'])" 2>&1 | tail -1)"
done
$T --master-port 29543 bench.py --gpus 2 --steps 200 --no-cpu-baseline --no-view-sharded --no-train-step --skip-e2e --grad-allreduce nccl > $O/r2m_n2_nccl.json 2> $O/r2m_n2_nccl.err
ts n2-nccl "rc=$? $(python -c "import json;d=json.loads(open('$O/r2m_n2_nccl.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"
$T --master-port 29544 bench.py --gpus 2 --steps 200 --no-cpu-baseline --no-view-sharded --no-train-step --skip-e2e --no-grad-allreduce > $O/r2m_n2_noar.json 2> $O/r2m_n2_noar.err
ts n2-no-allreduce "rc=$? $(python -c "import json;d=json.loads(open('$O/r2m_n2_noar.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"
tail -8 $O/r2m_n2_peer1.err > $O/r2m_n2_peer1_tail.txt
timeout 200 python tools/rows_gemm_timeline.py > $O/r2m_rows_gemm_timeline.txt 2>&1
ts rows-gemm-timeline "$(head -3 $O/r2m_rows_gemm_timeline.txt | tr '\n' ' ')"
