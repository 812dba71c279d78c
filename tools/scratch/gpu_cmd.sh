timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/bench_r1_n2.json 2>gpurun_out/bench_r1_n2.err
tail -2 gpurun_out/bench_r1_n2.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_r1_n2.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e'])"
timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/check_view_sharded.py > gpurun_out/view_sharded_n2.json 2>gpurun_out/view_sharded_n2.err
tail -2 gpurun_out/view_sharded_n2.err; cat gpurun_out/view_sharded_n2.json | tail -3
