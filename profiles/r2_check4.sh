#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_page_sizes.py -m gpu -q -x 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_full_size.py -m gpu -q -x -k "q1 or reproducible or acero" 2>&1 | tail -3
echo "== q1 sf10 / sf100"; timeout 300 python profiles/run_shape.py q1 59986052 4 2>&1 | tail -2; timeout 300 python profiles/run_shape.py q1 600037902 3 2>&1 | tail -1
echo "== q1d"; timeout 300 python profiles/run_shape.py q1d 59986052 3 2>&1 | tail -1
