#!/bin/bash
# GPU session G: repeatable comparison of stream-priority / SM-cap defaults (200 steps each).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/g_times.log; }
ts start
B="timeout 300 python bench.py --no-cpu-baseline --skip-e2e --steps 200"
run() { name=$1; shift; env "$@" $B > $O/g_bench_$name.json 2> $O/g_bench_$name.err; ts bench-$name "$(python -c "import json;d=json.load(open('$O/g_bench_$name.json'));print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"; }
for rep in 1 2; do
run default_$rep X=1
run all132_$rep SGC_TC_MAX_CTAS=132
run c112b1_$rep SGC_CHAIN_PRIO=-1,-1,-2 SGC_BIG_PRIO=x,x,-1
run c112b1_all132_$rep SGC_CHAIN_PRIO=-1,-1,-2 SGC_BIG_PRIO=x,x,-1 SGC_TC_MAX_CTAS=132
run inv_$rep SGC_CHAIN_PRIO=-3,-2,-1 SGC_BIG_PRIO=-3,-2,x SGC_TC_MAX_CTAS=132
run c112_$rep SGC_CHAIN_PRIO=-1,-1,-2
done
