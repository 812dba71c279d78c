#!/bin/bash
# Round 2, session X (1 GPU): timing-only upper bound of fusing the query chain (two GEMMs less per direction), SM caps, priorities.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/r2x_times.log; }
ts start
B="timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --no-view-sharded --no-train-step --skip-e2e --steps 300"
run() { name=$1; shift; env "$@" $B > $O/r2x_ab_$name.json 2> $O/r2x_ab_$name.err; ts ab-$name "$(python -c "import json;d=json.load(open('$O/r2x_ab_$name.json'));print(d['value'],d['ms_per_step'],d['gpu_launches_per_step'])" 2>&1 | tail -1)"; }
run def_1 SGC_X=1
run skipqo_1 SGC_EXP_SKIP_QO=1
run def_2 SGC_X=1
run skipqo_2 SGC_EXP_SKIP_QO=1
run cap116 SGC_TC_MAX_CTAS_FWD=116
run cap100 SGC_TC_MAX_CTAS_FWD=100
run wprio0 SGC_WSTREAM_PRIO=0
run def_3 SGC_X=1
