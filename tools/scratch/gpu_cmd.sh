timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 400 python bench.py > gpurun_out/b_full.json 2>gpurun_out/b_full.err
tail -2 gpurun_out/b_full.err
python -c "
import json;d=json.loads(open('gpurun_out/b_full.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','gpu_launches','roofline','path_roofline','cpu_baseline','clocks','kernels'): print(k, d.get(k))
"
