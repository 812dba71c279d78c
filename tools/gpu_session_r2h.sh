#!/bin/bash
# Round 2, session H: evidence for profiles/ -- full GPU suite, the default bench line, ncu launch list of the eager step,
# full-set captures of the dominant kernels (lift_bwd, lift_fwd, rows_gemm_tc, cross-view attention).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/r2h_times.log; }
ts start
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/r2h_suite.log
ts suite "$(tail -1 $O/r2h_suite.log)"
timeout 900 python bench.py > $O/r2h_bench_n1.json 2> $O/r2h_bench_n1.err
ts bench-default "$(python -c "import json;d=json.load(open('$O/r2h_bench_n1.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline'],d['loss_vs_oracle_rel'])" 2>&1 | tail -1)"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2h_bench_reference.json 2> $O/r2h_bench_reference.err
ts bench-reference "$(cut -c1-160 $O/r2h_bench_reference.json)"
CMD="python bench.py --no-graph --no-cpu-baseline --no-reference-gpu --no-view-sharded --no-train-step --skip-e2e --steps 3 --warmup 3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r2h_launches.csv $CMD > $O/r2h_ncu_launches.log 2>&1
ts launch-list "$(wc -l < $O/r2h_launches.csv)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lift_bwd_kernel -s 18 -c 3 -o $O/r2h_lift_bwd -f $CMD > $O/r2h_ncu_lift_bwd.log 2>&1
ts ncu-lift-bwd
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lift_fwd_kernel -s 18 -c 3 -o $O/r2h_lift_fwd -f $CMD > $O/r2h_ncu_lift_fwd.log 2>&1
ts ncu-lift-fwd
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rows_gemm_tc_kernel -s 252 -c 42 -o $O/r2h_rows_gemm -f $CMD > $O/r2h_ncu_rows_gemm.log 2>&1
ts ncu-rows-gemm
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:attn_(fwd|bwd_qt|bwd_slots)_kernel" -s 54 -c 9 -o $O/r2h_attn -f $CMD > $O/r2h_ncu_attn.log 2>&1
ts ncu-attn
timeout 300 python bench.py --no-cpu-baseline --no-reference-gpu --no-view-sharded --no-train-step --steps 50 --config SGCDet_large_ScanNet200 > $O/r2h_bench_large.json 2> $O/r2h_bench_large.err
ts bench-large "$(python -c "import json;d=json.load(open('$O/r2h_bench_large.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'])" 2>&1 | tail -1)"
ls -la $O/*.ncu-rep >> $O/r2h_times.log 2>&1
