#!/bin/bash
# A/B: stage C after the scan (PGF_PROBE_CONCURRENT=0) or concurrently with it on a second stream (default)
timeout 300 python -m pytest tests/test_gpu_join.py tests/test_gpu_sort.py -m gpu -q -x 2>&1 | tail -3
for sz in 59986052 600037902; do
  for conc in 0 1; do echo "== q3 rows=$sz concurrent=$conc"; PGF_PROBE_CONCURRENT=$conc Q3_LIMIT=10 timeout 120 python profiles/run_shape.py q3 $sz 4 2>&1 | tail -1 | cut -c1-175; done
done
