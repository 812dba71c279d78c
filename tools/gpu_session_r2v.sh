#!/bin/bash
# Round 2, session V (1 GPU): full suite (own Philox dropout masks, separable upsample backward); A/B of the separable backward,
# the lift_bwd prefetch and the CTA counts of the coarse levels' GEMMs; "-L" config; kernel table; timeline.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/r2v_times.log; }
ts start
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > $O/r2v_suite.log
ts suite "$(tail -1 $O/r2v_suite.log)"
B="timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --no-view-sharded --no-train-step --skip-e2e --steps 300"
run() { name=$1; shift; env "$@" $B > $O/r2v_ab_$name.json 2> $O/r2v_ab_$name.err; ts ab-$name "$(python -c "import json;d=json.load(open('$O/r2v_ab_$name.json'));print(d['value'],d['ms_per_step'],d['gpu_launches_per_step'])" 2>&1 | tail -1)"; }
run def_1 SGC_X=1
run sep0 SGC_UP_BWD_SEP=0
run pf0 SGC_LIFT_BWD_PF=0
run def_2 SGC_X=1
run small4 SGC_ROWS_SMALL_WORKS=4
run small8 SGC_ROWS_SMALL_WORKS=8
run small16 SGC_ROWS_SMALL_WORKS=16
run small8_t8 SGC_ROWS_SMALL_WORKS=8 SGC_ROWS_SMALL_TILES=8
run small16_t8 SGC_ROWS_SMALL_WORKS=16 SGC_ROWS_SMALL_TILES=8
run small8_cap140 SGC_ROWS_SMALL_WORKS=8 SGC_TC_MAX_CTAS_FWD=140
run small16_cap124 SGC_ROWS_SMALL_WORKS=16 SGC_TC_MAX_CTAS_FWD=124
run def_3 SGC_X=1
timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --no-view-sharded --no-train-step --steps 100 > $O/r2v_full.json 2> $O/r2v_full.err
ts full "$(python -c "import json;d=json.load(open('$O/r2v_full.json'));print(d['value'],d['ms_per_step'],d['roofline']['frac'],[(k,v['avg_ms']) for k,v in d['kernels'].items()])" 2>&1 | tail -1)"
L="timeout 300 python bench.py --config SGCDet_large_ScanNet200 --no-cpu-baseline --no-reference-gpu --no-view-sharded --no-train-step --steps 100"
runl() { name=$1; shift; env "$@" $L > $O/r2v_large_$name.json 2> $O/r2v_large_$name.err; ts large-$name "$(python -c "import json;d=json.load(open('$O/r2v_large_$name.json'));print(d['value'],d['ms_per_step'],[(k,v['avg_ms']) for k,v in d['kernels'].items()])" 2>&1 | tail -1)"; }
runl def SGC_X=1
runl sep0 SGC_UP_BWD_SEP=0
runl pf0 SGC_LIFT_BWD_PF=0
SGC_GRAPH_TRACE=$O/r2v_trace.json timeout 300 python tools/profile_step.py > $O/r2v_profile.txt 2>&1
python tools/graph_timeline.py $O/r2v_trace.json 30 $O/r2v_timeline_all.txt > $O/r2v_timeline.txt 2>&1
rm -f $O/r2v_trace.json
ts timeline "$(head -1 $O/r2v_timeline.txt)"
