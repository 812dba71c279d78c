#!/bin/bash
# GPU session L: many-CTA top-k for the large levels: parity, large-config bench and timeline.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/l_times.log; }
ts start
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > $O/l_tests.log
ts full-tests "$(tail -1 $O/l_tests.log)"
timeout 300 python bench.py --no-cpu-baseline --steps 50 --config SGCDet_large_ScanNet200 > $O/r1e_bench_large.json 2> $O/l_bench_large.err
ts bench-large "$(python -c "import json;d=json.load(open('$O/r1e_bench_large.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['path_roofline']['frac_of_hbm'])" 2>&1 | tail -1)"
SGC_TOPK_MC_MIN=100000000 timeout 300 python bench.py --no-cpu-baseline --skip-e2e --steps 50 --config SGCDet_large_ScanNet200 > $O/l_bench_large_oldtopk.json 2> $O/l_bench_large_oldtopk.err
ts bench-large-oldtopk "$(python -c "import json;d=json.load(open('$O/l_bench_large_oldtopk.json'));print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"
timeout 300 python bench.py --no-cpu-baseline --skip-e2e --steps 200 > $O/l_bench_default.json 2> $O/l_bench_default.err
ts bench-default "$(python -c "import json;d=json.load(open('$O/l_bench_default.json'));print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"
SGC_GRAPH_TRACE=$O/l_trace.json timeout 300 python tools/profile_step.py SGCDet_large_ScanNet200 40 > $O/l_profile_step_large.txt 2>&1
python tools/graph_timeline.py $O/l_trace.json 40 $O/l_timeline_large_all.txt > $O/l_timeline_large.txt 2>&1
rm -f $O/l_trace.json
ts timeline
