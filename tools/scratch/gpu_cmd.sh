timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for v in "A=1" "A=2"; do
  env $v timeout 200 python bench.py --no-cpu-baseline --skip-e2e > gpurun_out/b_tmp.json 2>gpurun_out/b_tmp.err
  echo "$v: $(python -c "import json;d=json.loads(open('gpurun_out/b_tmp.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['gpu_launches'])" 2>&1 | tail -1)"
done
tail -3 gpurun_out/b_tmp.err
