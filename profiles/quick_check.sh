python -m pytest tests/test_gpu_full_size.py -x -q 2>&1 | tail -15
