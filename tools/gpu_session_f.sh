#!/bin/bash
# GPU session F: wider converter stage in the weight-gradient kernels, stream priority variants.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/f_times.log; }
ts start
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > $O/f_tests.log
ts full-tests "$(tail -1 $O/f_tests.log)"
B="timeout 300 python bench.py --no-cpu-baseline --skip-e2e --steps 100"
run() { name=$1; shift; env "$@" $B > $O/f_bench_$name.json 2> $O/f_bench_$name.err; ts bench-$name "$(python -c "import json;d=json.load(open('$O/f_bench_$name.json'));print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"; }
run default X=1
run chain321 SGC_CHAIN_PRIO=-1,-2,-3
run big2 SGC_BIG_PRIO=x,x,-1
run chain321_big2 SGC_CHAIN_PRIO=-1,-2,-3 SGC_BIG_PRIO=x,x,-2
run chain321_big3 SGC_CHAIN_PRIO=-1,-2,-3 SGC_BIG_PRIO=x,x,-3
run chain111_big2 SGC_CHAIN_PRIO=-1,-1,-2 SGC_BIG_PRIO=x,x,-1
run bigall SGC_BIG_PRIO=0,0,-1
run chain123 SGC_CHAIN_PRIO=-3,-2,-1
SGC_CHAIN_PRIO=-1,-2,-3 SGC_BIG_PRIO=x,x,-2 SGC_GRAPH_TRACE=$O/f_trace.json timeout 300 python tools/profile_step.py > $O/f_profile_step.txt 2>&1
python tools/graph_timeline.py $O/f_trace.json 20 $O/f_timeline_all.txt > $O/f_timeline.txt 2>&1
rm -f $O/f_trace.json
ts timeline
