#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_join.py tests/test_gpu_sort.py tests/test_gpu_pipeline.py -m gpu -q -x 2>&1 | tail -2
for sz in 59986052 600037902; do echo "== q3 rows=$sz"; Q3_LIMIT=10 timeout 120 python profiles/run_shape.py q3 $sz 4 2>&1 | tail -1 | cut -c1-175; done
