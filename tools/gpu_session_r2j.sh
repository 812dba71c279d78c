#!/bin/bash
# Round 2, session J (1 GPU): depth-producer kernels vs the oracle, peer tests incl. the two-shot all-reduce, full bench line.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/r2j_times.log; }
ts start
timeout 600 python -m pytest tests/test_gpu_depth.py tests/test_gpu_peer.py -q 2>&1 | tail -40 > $O/r2j_new_tests.log
ts new-tests "$(tail -1 $O/r2j_new_tests.log)"
timeout 900 python bench.py > $O/r2j_bench_n1.json 2> $O/r2j_bench_n1.err
ts bench-default "$(python -c "import json;d=json.load(open('$O/r2j_bench_n1.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['depth_producer_bench'],d['view_sharded'])" 2>&1 | tail -1)"
tail -5 $O/r2j_bench_n1.err > $O/r2j_bench_n1_tail.txt
