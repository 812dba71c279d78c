#!/bin/bash
# ncu launch list of the eager step + full-set captures of the new tensor-core kernels.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/prof_times.log; }
ts start
timeout 300 python -m pytest tests/test_gpu_rows_gemm.py tests/test_gpu_path.py -x -q 2>&1 | tail -5 > $O/prof_tests.log
ts tests "$(tail -1 $O/prof_tests.log)"
timeout 300 python bench.py --no-cpu-baseline --skip-e2e --steps 100 > $O/prof_bench_default.json 2> $O/prof_bench_default.err
ts bench "$(python -c "import json;d=json.load(open('$O/prof_bench_default.json'));print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"
CMD="python bench.py --no-graph --no-cpu-baseline --skip-e2e --steps 3 --warmup 3"
# 3 eager warm-up + 1 recorder + 3 warm-up steps precede the timed ones: ~250 launches per step -> skip 1600, take 3 steps
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1600 -c 780 --csv --log-file $O/launches_r1c.csv $CMD > $O/prof_ncu_launches.log 2>&1
ts launch-list
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rows_gemm_tc -s 182 -c 4 -o $O/prof_rows_gemm_r1 -f $CMD > $O/prof_ncu_rows_gemm.log 2>&1
ts ncu-rows-gemm
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rows_wgrad_tc_kernel -s 84 -c 3 -o $O/prof_rows_wgrad_r1 -f $CMD > $O/prof_ncu_rows_wgrad.log 2>&1
ts ncu-rows-wgrad
