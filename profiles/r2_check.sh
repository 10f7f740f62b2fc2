#!/bin/bash
# Round 2: parity suite + Q3 kernel times + the bench end to end (1 GPU).
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2_tests5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests5.log
tail -12 gpurun_out/r2_tests5.log
for sz in 59986052 600037902; do echo "== q3 rows=$sz"; timeout 300 python profiles/run_shape.py q3 $sz 3 2>&1 | tail -1; done
echo "== q6 / q1 sf10"; timeout 300 python profiles/run_shape.py q6 59986052 3 2>&1 | tail -1; timeout 300 python profiles/run_shape.py q1 59986052 3 2>&1 | tail -1
echo "== bench ours"; timeout 1500 python bench.py --steps 5 --warmup 3 --record-expected > gpurun_out/r2_bench_try.json 2> gpurun_out/r2_bench_try.err; tail -c 600 gpurun_out/r2_bench_try.json; tail -5 gpurun_out/r2_bench_try.err
cp profiles/sf_expected.json gpurun_out/sf_expected.json 2>/dev/null
