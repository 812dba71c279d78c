#!/bin/bash
# GPU session B: per-level chain streams, forward SM caps, rows-GEMM column split variants.  Outputs in gpurun_out/.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/b_times.log; }
ts start
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > $O/b_tests.log
ts full-tests
B="timeout 300 python bench.py --no-cpu-baseline --skip-e2e"
run() { name=$1; shift; env "$@" $B > $O/b_bench_$name.json 2> $O/b_bench_$name.err; ts bench-$name "$(python -c "import json;d=json.load(open('$O/b_bench_$name.json'));print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"; }
run default X=1
run lvl0 SGC_LEVEL_STREAMS=0
run fwd132 SGC_TC_MAX_CTAS_FWD=132
run fwd116 SGC_TC_MAX_CTAS_FWD=116
run fwd100 SGC_TC_MAX_CTAS_FWD=100
run fwd116_all140 SGC_TC_MAX_CTAS_FWD=116 SGC_TC_MAX_CTAS=140
run fwd116_ncta256 SGC_TC_MAX_CTAS_FWD=116 SGC_ROWS_NCTA=256
run fwd116_ncta128 SGC_TC_MAX_CTAS_FWD=116 SGC_ROWS_NCTA=128
run fwd116_noprezero SGC_TC_MAX_CTAS_FWD=116 SGC_PREZERO=0
SGC_TC_MAX_CTAS_FWD=116 SGC_GRAPH_TRACE=$O/b_trace.json timeout 300 python tools/profile_step.py > $O/b_profile_step.txt 2>&1
python tools/graph_timeline.py $O/b_trace.json 20 $O/b_timeline_all.txt > $O/b_timeline.txt 2>&1
rm -f $O/b_trace.json
ts timeline
