#!/bin/bash
# Round 2: parity suite, Q3 kernel times for tile sizes / consumer-warp counts, the new bench end to end, and one
# ncu --set full capture of the three Q3 pipelines at SF100.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2_tests4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests4.log
tail -12 gpurun_out/r2_tests4.log
for sz in 59986052 600037902; do
  echo "== default(20 warps) rows=$sz"; timeout 300 python profiles/run_shape.py q3 $sz 3 2>&1 | tail -1
  for w in 16 24; do
    echo "== variant w$w rows=$sz"; PGF_B200_LIB=$PWD/pg_fusion_b200/variants/libpgf_b200_w$w.so timeout 300 python profiles/run_shape.py q3 $sz 3 2>&1 | tail -1
  done
done
echo "== q3var sf10"; timeout 300 python profiles/run_shape.py q3var 59986052 3 2>&1 | tail -5
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2_bench_ref_try.json 2> gpurun_out/r2_bench_ref_try.err; tail -c 1500 gpurun_out/r2_bench_ref_try.json; tail -3 gpurun_out/r2_bench_ref_try.err
echo "== bench ours"; timeout 1500 python bench.py --steps 3 --warmup 2 --record-expected > gpurun_out/r2_bench_try.json 2> gpurun_out/r2_bench_try.err; tail -c 3000 gpurun_out/r2_bench_try.json; tail -5 gpurun_out/r2_bench_try.err
cp profiles/sf_expected.json gpurun_out/sf_expected.json 2>/dev/null
python profiles/run_shape.py q3 600037902 3 > gpurun_out/r2_q3_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:probe_pipeline -s 3 -c 3 -o gpurun_out/r2_q3_sf100 python profiles/run_shape.py q3 600037902 3 > gpurun_out/r2_q3_ncu.log 2>&1
tail -2 gpurun_out/r2_q3_ncu.log
