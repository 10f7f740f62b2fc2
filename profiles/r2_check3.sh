#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_join.py tests/test_gpu_cpp_host.py -m gpu -q -x 2>&1 | tail -15
echo "== q3 sf100"; Q3_LIMIT=10 timeout 300 python profiles/run_shape.py q3 600037902 4 2>&1 | tail -1
echo "== q3 sf10"; Q3_LIMIT=10 timeout 300 python profiles/run_shape.py q3 59986052 4 2>&1 | tail -1
echo "== q6 q1 sf10"; timeout 300 python profiles/run_shape.py q6 59986052 3 2>&1 | tail -1; timeout 300 python profiles/run_shape.py q1 59986052 3 2>&1 | tail -1
