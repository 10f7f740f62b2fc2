#!/bin/bash
# Round 2, first GPU pass of the compaction pipeline: parity tests, then Q3 kernel times at SF10 / SF100
# for the default build and the consumer-warp-count variants (pg_fusion_b200/variants/, built with
# make BUILD=build_wN OUT=../variants/libpgf_b200_wN.so EXTRA=-DPGF_PROBE_WARPS=N).
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt; nvidia-smi -L >> gpurun_out/host.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests1.log
tail -25 gpurun_out/r2_tests1.log
for sz in 59986052 600037902; do
  echo "== default (20 warps) rows=$sz"; timeout 300 python profiles/run_shape.py q3 $sz 4 2>&1 | tail -4
done
for w in 16 24 28; do
  for sz in 59986052 600037902; do
    echo "== variant w$w rows=$sz"; PGF_B200_LIB=$PWD/pg_fusion_b200/variants/libpgf_b200_w$w.so timeout 300 python profiles/run_shape.py q3 $sz 4 2>&1 | tail -2
  done
done
echo "== q3var sf10"; timeout 300 python profiles/run_shape.py q3var 59986052 3 2>&1 | tail -5
echo "== q3bloom sf10"; timeout 300 python profiles/run_shape.py q3bloom 59986052 3 2>&1 | tail -2
