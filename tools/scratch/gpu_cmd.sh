timeout 600 python -m pytest tests/test_gpu_rowops.py tests/test_gpu_path.py -q -x 2>&1 | tail -5
for v in "SGC_SMALL_ROWS=0" "SGC_SMALL_ROWS=1000"; do
  env $v timeout 200 python bench.py --no-cpu-baseline --skip-e2e > gpurun_out/b_tmp.json 2>gpurun_out/b_tmp.err
  echo "$v: $(python -c "import json;d=json.loads(open('gpurun_out/b_tmp.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['gpu_launches'])" 2>&1 | tail -1)"
done
tail -3 gpurun_out/b_tmp.err
SGC_GRAPH_TRACE=gpurun_out/graph_trace8.json timeout 200 python tools/profile_step.py > gpurun_out/prof_step2.txt 2>&1
tail -2 gpurun_out/prof_step2.txt
