#!/bin/bash
# GPU session H: output_proj x query-projection fusion on the chain, wait_group.read at the GEMM kernel's exit.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/h_times.log; }
ts start
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > $O/h_tests.log
ts full-tests "$(tail -1 $O/h_tests.log)"
B="timeout 300 python bench.py --no-cpu-baseline --skip-e2e --steps 200"
run() { name=$1; shift; env "$@" $B > $O/h_bench_$name.json 2> $O/h_bench_$name.err; ts bench-$name "$(python -c "import json;d=json.load(open('$O/h_bench_$name.json'));print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"; }
for rep in 1 2 3; do
run default_$rep X=1
run noqo_$rep SGC_FUSE_QO=0
done
