#!/bin/bash
# Round 2 A/B: late-column prefetch at the top of stage C (off / L1 / L2), new top-k.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sort.py tests/test_gpu_join.py -m gpu -q -x 2>&1 | tail -2
for sz in 59986052 600037902; do
  for v in nopf default pfl2; do
    echo "== q3 rows=$sz $v"
    if [ $v = default ]; then unset PGF_B200_LIB; else export PGF_B200_LIB=$PWD/pg_fusion_b200/variants/libpgf_b200_$v.so; fi
    Q3_LIMIT=10 timeout 300 python profiles/run_shape.py q3 $sz 4 2>&1 | tail -2
  done
done
