#!/bin/bash
# Round 2, session AD (1 GPU), final state: smoke, full suite, the default bench line (all legs), the reference arm, ncu launch
# list of the eager step, full-set captures of the lift kernels (summarised on the box), "-L" and V=100 lines, timeline.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/r2ad_times.log; }
ts start
timeout 300 python __graft_entry__.py smoke > $O/r2ad_smoke.log 2>&1
ts smoke "$(tail -1 $O/r2ad_smoke.log)"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/r2ad_suite.log
ts suite "$(tail -1 $O/r2ad_suite.log)"
timeout 900 python bench.py > $O/r2ad_bench_n1.json 2> $O/r2ad_bench_n1.err
ts bench-default "$(python -c "import json;d=json.load(open('$O/r2ad_bench_n1.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline'],d['loss_vs_oracle_rel'],d['gpu_launches_per_step'])" 2>&1 | tail -1)"
timeout 300 python bench.py --steps 300 --no-cpu-baseline --no-reference-gpu --no-view-sharded --no-train-step --skip-e2e > $O/r2ad_bench_n1_300.json 2> $O/r2ad_bench_n1_300.err
ts bench-300 "$(python -c "import json;d=json.load(open('$O/r2ad_bench_n1_300.json'));print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2ad_bench_reference.json 2> $O/r2ad_bench_reference.err
ts bench-reference "$(cut -c1-160 $O/r2ad_bench_reference.json)"
CMD="python bench.py --no-graph --no-cpu-baseline --no-reference-gpu --no-view-sharded --no-train-step --skip-e2e --steps 3 --warmup 3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r2ad_launches.csv $CMD > $O/r2ad_ncu_launches.log 2>&1
ts launch-list "$(wc -l < $O/r2ad_launches.csv)"
python tools/summarize_ncu.py launches $O/r2ad_launches.csv $O/r2ad_launches_summary.md > /dev/null 2>&1
cap() { # name regex skip count
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -o $O/r2ad_$1 -f $CMD > $O/r2ad_ncu_$1.log 2>&1
  python tools/summarize_ncu.py full $O/r2ad_$1.ncu-rep $O/r2ad_$1_full_summary.md $O/r2ad_$1_traffic.json > /dev/null 2>&1
  ts ncu-$1 "$(ls -la $O/r2ad_$1.ncu-rep | awk '{print $5}')"
}
cap lift_bwd lift_bwd_kernel 18 3
cap lift_fwd lift_fwd_kernel 18 3
timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --no-view-sharded --no-train-step --steps 50 --config SGCDet_large_ScanNet200 > $O/r2ad_bench_large.json 2> $O/r2ad_bench_large.err
ts bench-large "$(python -c "import json;d=json.load(open('$O/r2ad_bench_large.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'])" 2>&1 | tail -1)"
timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --no-view-sharded --no-train-step --steps 50 --views 100 > $O/r2ad_bench_v100.json 2> $O/r2ad_bench_v100.err
ts bench-v100 "$(python -c "import json;d=json.load(open('$O/r2ad_bench_v100.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'])" 2>&1 | tail -1)"
SGC_GRAPH_TRACE=$O/r2ad_trace.json timeout 300 python tools/profile_step.py > $O/r2ad_profile.txt 2>&1
python tools/graph_timeline.py $O/r2ad_trace.json 30 $O/r2ad_timeline_all.txt > $O/r2ad_timeline.txt 2>&1
rm -f $O/r2ad_trace.json
ts timeline "$(head -1 $O/r2ad_timeline.txt)"
while [ "$(du -sm $O | cut -f1)" -gt 56 ]; do f=$(ls -S $O/*.ncu-rep 2>/dev/null | head -1); [ -z "$f" ] && break; rm -f "$f"; ts dropped $f; done
