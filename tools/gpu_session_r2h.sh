#!/bin/bash
# Round 2, session H: peer-memory collectives + view-sharded product path (1 GPU: simulated ranks, two processes sharing the
# GPU), full GPU suite, the default bench line, ncu launch list of the eager step, full-set captures of the dominant kernels
# (summarised on the box: the .ncu-rep files stay below gpurun's 64 MiB return limit).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/r2h_times.log; }
ts start
timeout 600 python -m pytest tests/test_gpu_peer.py -x -q 2>&1 | tail -30 > $O/r2h_peer_tests.log
ts peer-tests "$(tail -1 $O/r2h_peer_tests.log)"
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_peer.py 2>&1 | tail -15 > $O/r2h_suite.log
ts suite "$(tail -1 $O/r2h_suite.log)"
timeout 900 python bench.py > $O/r2h_bench_n1.json 2> $O/r2h_bench_n1.err
ts bench-default "$(python -c "import json;d=json.load(open('$O/r2h_bench_n1.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline'],d['loss_vs_oracle_rel'],d['view_sharded'])" 2>&1 | tail -1)"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2h_bench_reference.json 2> $O/r2h_bench_reference.err
ts bench-reference "$(cut -c1-160 $O/r2h_bench_reference.json)"
CMD="python bench.py --no-graph --no-cpu-baseline --no-reference-gpu --no-view-sharded --no-train-step --skip-e2e --steps 3 --warmup 3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r2h_launches.csv $CMD > $O/r2h_ncu_launches.log 2>&1
ts launch-list "$(wc -l < $O/r2h_launches.csv)"
cap() { # name regex skip count
  timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -o $O/r2h_$1 -f $CMD > $O/r2h_ncu_$1.log 2>&1
  python tools/summarize_ncu.py full $O/r2h_$1.ncu-rep $O/r2h_$1_full_summary.md $O/r2h_$1_traffic.json > /dev/null 2>&1
  ts ncu-$1 "$(ls -la $O/r2h_$1.ncu-rep | awk '{print $5}')"
}
cap lift_bwd lift_bwd_kernel 18 3
cap lift_fwd lift_fwd_kernel 18 3
cap rows_gemm rows_gemm_tc_kernel 252 8
cap attn "attn_(fwd|bwd_qt|bwd_slots)_kernel" 54 9
timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --no-view-sharded --no-train-step --steps 50 --config SGCDet_large_ScanNet200 > $O/r2h_bench_large.json 2> $O/r2h_bench_large.err
ts bench-large "$(python -c "import json;d=json.load(open('$O/r2h_bench_large.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'])" 2>&1 | tail -1)"
# gpurun returns at most 64 MiB: drop the largest reports first if needed (their summaries stay)
while [ "$(du -sm $O | cut -f1)" -gt 56 ]; do f=$(ls -S $O/*.ncu-rep 2>/dev/null | head -1); [ -z "$f" ] && break; rm -f "$f"; ts dropped $f; done
