#!/bin/bash
# final measurements of the round (bench lines for profiles/, smoke, timeline).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/meas_times.log; }
ts start
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/meas_tests.log
ts full-tests "$(tail -1 $O/meas_tests.log)"
timeout 300 python __graft_entry__.py smoke > $O/meas_smoke.log 2>&1
ts smoke "$(tail -1 $O/meas_smoke.log)"
timeout 400 python bench.py --steps 100 > $O/r1x_bench_n1.json 2> $O/meas_bench_n1.err
ts bench-n1 "$(python -c "import json;d=json.load(open('$O/r1x_bench_n1.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['kernel'],d['roofline']['frac'])" 2>&1 | tail -1)"
timeout 400 python bench.py > $O/r1x_bench_n1_default_args.json 2> $O/meas_bench_n1d.err
ts bench-n1-default-args "$(python -c "import json;d=json.load(open('$O/r1x_bench_n1_default_args.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'])" 2>&1 | tail -1)"
timeout 300 python bench.py --impl reference > $O/r1x_bench_reference.json 2> $O/meas_bench_ref.err
ts bench-reference "$(cut -c1-120 $O/r1x_bench_reference.json)"
timeout 300 python bench.py --no-cpu-baseline --steps 50 --views 100 > $O/r1x_bench_v100.json 2> $O/meas_bench_v100.err
ts bench-v100 "$(python -c "import json;d=json.load(open('$O/r1x_bench_v100.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['path_roofline']['frac_of_hbm'])" 2>&1 | tail -1)"
timeout 300 python bench.py --no-cpu-baseline --steps 50 --config SGCDet_large_ScanNet200 > $O/r1x_bench_large.json 2> $O/meas_bench_large.err
ts bench-large "$(python -c "import json;d=json.load(open('$O/r1x_bench_large.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['path_roofline']['frac_of_hbm'])" 2>&1 | tail -1)"
timeout 300 python bench.py --no-cpu-baseline --steps 50 --scenes-per-gpu 4 > $O/r1x_bench_n1_b4.json 2> $O/meas_bench_b4.err
ts bench-b4 "$(python -c "import json;d=json.load(open('$O/r1x_bench_n1_b4.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'])" 2>&1 | tail -1)"
SGC_GRAPH_TRACE=$O/meas_trace.json timeout 300 python tools/profile_step.py > $O/meas_profile_step.txt 2>&1
python tools/graph_timeline.py $O/meas_trace.json 30 $O/meas_timeline_all.txt > $O/r1x_graph_timeline.txt 2>&1
rm -f $O/meas_trace.json
ts timeline
