#!/bin/bash
# GPU-side capture of one round's evidence: usage  bash profiles/capture.sh <tag>
# (run under gpurun; outputs land in gpurun_out/ and are summarised locally by profiles/summarize.py)
set -x
TAG=${1:-r1}
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || exit 1
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/ncu_launch_$TAG.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
# dominant kernel of the headline workload (Q6 SF10): second launch
$NCU -k regex:pipeline_kernel -s 1 -c 1 -f -o gpurun_out/prof_q6_$TAG python profiles/run_shape.py q6 59986052 3 > gpurun_out/ncu_q6_$TAG.log 2>&1
$NCU -k regex:pipeline_kernel -s 1 -c 1 -f -o gpurun_out/prof_q1_$TAG python profiles/run_shape.py q1 59986052 3 > gpurun_out/ncu_q1_$TAG.log 2>&1
$NCU -k regex:pipeline_kernel -s 5 -c 1 -f -o gpurun_out/prof_q3_$TAG python profiles/run_shape.py q3 59986052 3 > gpurun_out/ncu_q3_$TAG.log 2>&1
$NCU -k regex:pipeline_kernel -s 5 -c 1 -f -o gpurun_out/prof_q3bloom_$TAG python profiles/run_shape.py q3bloom 59986052 3 > gpurun_out/ncu_q3bloom_$TAG.log 2>&1
$NCU -k regex:bloom_probe -s 2 -c 1 -f -o gpurun_out/prof_bloom_$TAG python profiles/run_shape.py bloom 64000000 4 > gpurun_out/ncu_bloom_$TAG.log 2>&1
cat gpurun_out/bench_$TAG.json
