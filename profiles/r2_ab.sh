#!/bin/bash
# Round 2 A/B: stage A with one row (variant a1) or two rows (default) per lane and step.
timeout 900 python -m pytest tests/test_gpu_join.py tests/test_gpu_exchange.py tests/test_gpu_bloom.py -m gpu -q -x 2>&1 | tail -2
for sz in 59986052 600037902; do
  for v in a1 default; do
    echo "== q3 rows=$sz $v"
    if [ $v = default ]; then unset PGF_B200_LIB; else export PGF_B200_LIB=$PWD/pg_fusion_b200/variants/libpgf_b200_$v.so; fi
    Q3_LIMIT=10 timeout 300 python profiles/run_shape.py q3 $sz 4 2>&1 | tail -2
  done
done
unset PGF_B200_LIB
timeout 600 python -m pytest tests/test_gpu_full_size.py -m gpu -q -x -k "q3" 2>&1 | tail -2
