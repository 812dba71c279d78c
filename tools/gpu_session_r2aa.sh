#!/bin/bash
# Round 2, session AA (1 GPU): fused query chain (output_proj / query projection / key product as one GEMM): tests, A/B, timeline.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/r2aa_times.log; }
ts start
timeout 300 python -m pytest tests/test_gpu_rows_gemm.py -q -x -k fused_query 2>&1 | tail -15 > $O/r2aa_fused.log
ts fused-test "$(tail -1 $O/r2aa_fused.log)"
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > $O/r2aa_suite.log
ts suite "$(tail -1 $O/r2aa_suite.log)"
B="timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --no-view-sharded --no-train-step --skip-e2e --steps 300"
run() { name=$1; shift; env "$@" $B > $O/r2aa_ab_$name.json 2> $O/r2aa_ab_$name.err; ts ab-$name "$(python -c "import json;d=json.load(open('$O/r2aa_ab_$name.json'));print(d['value'],d['ms_per_step'],d['gpu_launches_per_step'],d['loss'])" 2>&1 | tail -1)"; }
run def_1 SGC_X=1
run unfused_1 SGC_FUSE_QUERY=0
run def_2 SGC_X=1
run unfused_2 SGC_FUSE_QUERY=0
run def_3 SGC_X=1
run unfused_3 SGC_FUSE_QUERY=0
timeout 300 python bench.py --no-reference-gpu --no-view-sharded --no-train-step --steps 50 > $O/r2aa_full.json 2> $O/r2aa_full.err
ts full "$(python -c "import json;d=json.load(open('$O/r2aa_full.json'));print(d['value'],d['ms_per_step'],d['loss'],d['loss_vs_oracle_rel'],d['e2e']['value'])" 2>&1 | tail -1)"
SGC_GRAPH_TRACE=$O/r2aa_trace.json timeout 300 python tools/profile_step.py > $O/r2aa_profile.txt 2>&1
python tools/graph_timeline.py $O/r2aa_trace.json 30 $O/r2aa_timeline_all.txt > $O/r2aa_timeline.txt 2>&1
rm -f $O/r2aa_trace.json
ts timeline "$(head -1 $O/r2aa_timeline.txt)"
