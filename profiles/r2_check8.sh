#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_bloom.py tests/test_gpu_join.py tests/test_gpu_cpp_host.py tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -12
