#!/bin/bash
# Round 2, session G (2 GPUs): NCCL inside CUDA graphs (gradient all-reduce, view-sharded step), config-3 / config-5 legs.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/r2g_times.log; }
ts start
timeout 300 python -m pytest tests/test_gpu_lift_tiles.py tests/test_gpu_path.py -x -q -k 'tile or backward' 2>&1 | tail -3 > $O/r2g_tests.log
ts tests "$(tail -1 $O/r2g_tests.log)"
timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --no-view-sharded --no-train-step --skip-e2e --steps 200 > $O/r2g_bench_n1.json 2> $O/r2g_bench_n1.err
ts bench-n1 "$(python -c "import json;d=json.load(open('$O/r2g_bench_n1.json'));print(d['value'],d['ms_per_step'],d['gpu_launches_per_step'])" 2>&1 | tail -1)"
T="timeout -k 5 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$T --master-port 29511 bench.py --gpus 2 --steps 100 --no-cpu-baseline --no-view-sharded --no-train-step --skip-e2e > $O/r2g_n2_outside.json 2> $O/r2g_n2_outside.err
ts n2-allreduce-outside "$(python -c "import json;d=json.loads(open('$O/r2g_n2_outside.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"
$T --master-port 29512 bench.py --gpus 2 --steps 100 --no-cpu-baseline --no-view-sharded --no-train-step --skip-e2e --allreduce-in-graph 1 > $O/r2g_n2_ingraph.json 2> $O/r2g_n2_ingraph.err
ts n2-allreduce-in-graph "rc=$? $(python -c "import json;d=json.loads(open('$O/r2g_n2_ingraph.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"
$T --master-port 29513 bench.py --gpus 2 --steps 20 --no-cpu-baseline --skip-e2e --view-sharded-graph 1 > $O/r2g_n2_legs.json 2> $O/r2g_n2_legs.err
ts n2-legs "rc=$? $(python -c "import json;d=json.loads(open('$O/r2g_n2_legs.json').read().strip().splitlines()[-1]);print(d['value'],d['view_sharded'],d['train_step'])" 2>&1 | tail -1)"
tail -5 $O/r2g_n2_ingraph.err > $O/r2g_n2_ingraph_tail.txt; tail -5 $O/r2g_n2_legs.err > $O/r2g_n2_legs_tail.txt
