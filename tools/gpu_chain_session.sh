#!/bin/bash
# First GPU session of the next round: validate + bench the row-tile chain kernels (csrc/sgc_rows_chain*_tc.cu).
#   gpurun --timeout 600 -- 'bash tools/gpu_chain_session.sh'      -> gpurun_out/chain_*.{log,json}
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $O/chain_times.log; }
ts start
# the backward kernel has never run: its own test first, under a short timeout (a pipeline bug would hang, not fail)
SGC_TEST_CHAIN_BWD=1 timeout 60 python -m pytest tests/test_gpu_rows_chain.py -x -q 2>&1 | tail -15 > $O/chain_bwd_kernel.log
ts bwd-kernel-tests "$(tail -1 $O/chain_bwd_kernel.log)"
SGC_ROWS_CHAIN=1 timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/chain_fwd_suite.log
ts suite-with-forward-chain "$(tail -1 $O/chain_fwd_suite.log)"
SGC_ROWS_CHAIN=1 SGC_ROWS_CHAIN_BWD=1 timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/chain_both_suite.log
ts suite-with-both-chains "$(tail -1 $O/chain_both_suite.log)"
B="timeout 300 python bench.py --no-cpu-baseline --skip-e2e --steps 200"
run() { name=$1; shift; env "$@" $B > $O/chain_bench_$name.json 2> $O/chain_bench_$name.err; ts bench-$name "$(python -c "import json;d=json.load(open('$O/chain_bench_$name.json'));print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"; }
for rep in 1 2; do
run base_$rep X=1
run fwd_$rep SGC_ROWS_CHAIN=1
run both_$rep SGC_ROWS_CHAIN=1 SGC_ROWS_CHAIN_BWD=1
done
