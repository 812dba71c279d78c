set -x
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 700 -c 900 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --no-graph --steps 2 --warmup 3 --no-cpu-baseline --skip-e2e > gpurun_out/ncu_bench_b.log 2>&1
tail -c 300 gpurun_out/ncu_bench_b.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'lift_bwd|lift_fwd|project_tc_kernel|wgrad_tc_kernel|attn_fwd|attn_bwd_qt|rowop_fwd' -c 30 -o gpurun_out/prof_r1b_top python bench.py --no-graph --steps 1 --warmup 3 --no-cpu-baseline --skip-e2e > gpurun_out/ncu_full_b.log 2>&1
tail -c 300 gpurun_out/ncu_full_b.log
ls -la gpurun_out/*.ncu-rep
