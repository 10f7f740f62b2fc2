#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_cpp_host.py -m gpu -q -x 2>&1 | tail -8
