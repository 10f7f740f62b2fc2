#!/bin/bash
# Round 2: parity suite, then Q3 kernel times at SF10 / SF100 for tile sizes, ring depths and consumer-warp counts
# (variants in pg_fusion_b200/variants/: make BUILD=build_wN OUT=../variants/libpgf_b200_wN.so EXTRA=-DPGF_PROBE_WARPS=N),
# then one ncu --set full capture of the three Q3 pipelines at SF10.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2_tests3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests3.log
tail -25 gpurun_out/r2_tests3.log
for tr in 256 128 64; do for d in 4 2; do
  echo "== default(20 warps) tile_rows<=$tr depth<=$d"; PGF_PROBE_TILE_ROWS=$tr PGF_PROBE_DEPTH=$d timeout 300 python profiles/run_shape.py q3 59986052 3 2>&1 | tail -1
done; done
echo "== default SF100"; timeout 300 python profiles/run_shape.py q3 600037902 3 2>&1 | tail -1
for w in 16 24; do for tr in 256 128; do
  echo "== variant w$w tile_rows<=$tr"; PGF_PROBE_TILE_ROWS=$tr PGF_B200_LIB=$PWD/pg_fusion_b200/variants/libpgf_b200_w$w.so timeout 300 python profiles/run_shape.py q3 59986052 3 2>&1 | tail -1
done; done
echo "== q3var sf10"; timeout 300 python profiles/run_shape.py q3var 59986052 3 2>&1 | tail -5
echo "== q3bloom sf10"; timeout 300 python profiles/run_shape.py q3bloom 59986052 3 2>&1 | tail -1
python profiles/run_shape.py q3 59986052 3 > gpurun_out/r2_q3_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:probe_pipeline -s 3 -c 3 -o gpurun_out/r2_q3b python profiles/run_shape.py q3 59986052 3 > gpurun_out/r2_q3_ncu.log 2>&1
tail -2 gpurun_out/r2_q3_ncu.log
