#!/bin/bash
# Round 2, session U (1 GPU): full suite; A/B of the prepare-stream priorities (coarse levels' lift / projection gradients
# beside the finest level's lift backward), the wider lift_fwd gathers, small-problem CTA counts; kernel table; timeline.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/r2u_times.log; }
ts start
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > $O/r2u_suite.log
ts suite "$(tail -1 $O/r2u_suite.log)"
B="timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --no-view-sharded --no-train-step --skip-e2e --steps 300"
run() { name=$1; shift; env "$@" $B > $O/r2u_ab_$name.json 2> $O/r2u_ab_$name.err; ts ab-$name "$(python -c "import json;d=json.load(open('$O/r2u_ab_$name.json'));print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"; }
run base_1 SGC_PREP_PRIO=0,0,0 SGC_LIFT_FWD_MINB=4
run new_1 SGC_X=1
run prio_only SGC_LIFT_FWD_MINB=4
run lift_only SGC_PREP_PRIO=0,0,0
run base_2 SGC_PREP_PRIO=0,0,0 SGC_LIFT_FWD_MINB=4
run new_2 SGC_X=1
run small16 SGC_ROWS_SMALL_WORKS=16
run small16_cap116 SGC_ROWS_SMALL_WORKS=16 SGC_TC_MAX_CTAS_FWD=116
run small32_cap116 SGC_ROWS_SMALL_WORKS=32 SGC_TC_MAX_CTAS_FWD=116
run prio2 SGC_PREP_PRIO=-2,-2,0
run prio_all SGC_PREP_PRIO=-1,-1,-1
run fwd_minb2 SGC_LIFT_FWD_MINB=2
run bwd_minb3 SGC_LIFT_MINB=3
run base_3 SGC_PREP_PRIO=0,0,0 SGC_LIFT_FWD_MINB=4
run new_3 SGC_X=1
timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --no-view-sharded --no-train-step --steps 100 > $O/r2u_full.json 2> $O/r2u_full.err
ts full "$(python -c "import json;d=json.load(open('$O/r2u_full.json'));print(d['value'],d['ms_per_step'],d['roofline'],[(k,v['avg_ms']) for k,v in d['kernels'].items()])" 2>&1 | tail -1)"
SGC_GRAPH_TRACE=$O/r2u_trace.json timeout 300 python tools/profile_step.py > $O/r2u_profile.txt 2>&1
python tools/graph_timeline.py $O/r2u_trace.json 30 $O/r2u_timeline_all.txt > $O/r2u_timeline.txt 2>&1
rm -f $O/r2u_trace.json
ts timeline "$(head -1 $O/r2u_timeline.txt)"
