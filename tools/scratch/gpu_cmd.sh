timeout 600 python -m pytest tests/test_gpu_project_tc.py -q -x 2>&1 | tail -3
timeout 200 python tools/bench_project_tc.py 2>&1 | tail -5
timeout 200 python tools/bench_project_tc.py 40 128 256 2>&1 | tail -5
