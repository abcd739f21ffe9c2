timeout 1500 python -m pytest tests/test_gpu_large.py -x -q 2>&1 | tail -30
