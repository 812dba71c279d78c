timeout 600 python -m pytest tests/test_gpu_rowops.py tests/test_gpu_path.py -q 2>&1 | tail -15
