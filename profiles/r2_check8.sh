#!/bin/bash
# re-capture of the Q3 pipelines with the stage-C kernel included (see r2_capture.sh)
TAG=r2f
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k "regex:probe_pipeline|entries_pipeline" -s 4 -c 4 -f -o gpurun_out/prof_q3_sf10_$TAG python profiles/run_shape.py q3 59986052 3 > gpurun_out/ncu_q3_sf10_$TAG.log 2>&1; tail -1 gpurun_out/ncu_q3_sf10_$TAG.log
$NCU -k "regex:probe_pipeline|entries_pipeline" -s 4 -c 4 -f -o gpurun_out/prof_q3_sf100_$TAG python profiles/run_shape.py q3 600037902 3 > gpurun_out/ncu_q3_sf100_$TAG.log 2>&1; tail -1 gpurun_out/ncu_q3_sf100_$TAG.log
