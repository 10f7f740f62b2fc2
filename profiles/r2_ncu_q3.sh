#!/bin/bash
# Round 2: GPU parity suite, then one ncu --set full capture of the three Q3 pipelines (compaction kernel) at SF10.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests2.log
tail -15 gpurun_out/r2_tests2.log
python profiles/run_shape.py q3 59986052 3 > gpurun_out/r2_q3_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:probe_pipeline -s 3 -c 3 -o gpurun_out/r2_q3 python profiles/run_shape.py q3 59986052 3 > gpurun_out/r2_q3_ncu.log 2>&1
tail -4 gpurun_out/r2_q3_plain.log; tail -3 gpurun_out/r2_q3_ncu.log
python profiles/run_shape.py q3 600037902 3 2>&1 | tail -2
