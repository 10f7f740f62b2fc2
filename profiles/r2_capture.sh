#!/bin/bash
# Round 2 evidence, one GPU: parity suite, the bench (both arms), the ncu launch list of the bench command, and one
# ncu --set full capture per shape.  Outputs in gpurun_out/, summarised locally into profiles/ (read_ncu.py, hot_lines.py).
TAG=${1:-r2}   # the committed evidence of round 2 carries the tag r2f (final code)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -14 gpurun_out/${TAG}_tests.log
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_reference_arm.json 2> gpurun_out/bench_${TAG}_ref.err; tail -c 600 gpurun_out/bench_${TAG}_reference_arm.json
echo "== bench ours"; timeout 1500 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_${TAG}.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_${TAG}.json') if l.startswith('{')][-1])
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('h2d_GBps_per_gpu'), d['e2e'].get('h2d_link_peak_GBps'))
print({k:(round(v['ms_per_pass'],3), round(v['kernel_ms'],3), round(v['frac'],3)) for k,v in d['shapes'].items()})
print(d['shapes']['q3']['kernel_ms_by_pipeline'], d['shapes']['q3']['ms_per_pass_with_runtime_filters'], d['parity'].get('mismatches'), d['clocks'])
for k,v in d.get('other_workloads',{}).items(): print(k, {a:b for a,b in v.items() if 'frac' in a or a=='kernel_ms'})
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['all_cores']['value'], d['cpu_baseline']['all_cores']['cores'])
PY
echo "== launch list"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --sf 10 --steps 2 --warmup 3 --no-extras --e2e-steps 1 > gpurun_out/ncu_launch_${TAG}.log 2>&1; echo "rc=$?"
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:pipeline_kernel -s 1 -c 1 -f -o gpurun_out/prof_q6_$TAG python profiles/run_shape.py q6 600037902 3 > gpurun_out/ncu_q6_$TAG.log 2>&1; tail -1 gpurun_out/ncu_q6_$TAG.log
$NCU -k regex:pipeline_kernel -s 1 -c 1 -f -o gpurun_out/prof_q1_$TAG python profiles/run_shape.py q1 600037902 3 > gpurun_out/ncu_q1_$TAG.log 2>&1; tail -1 gpurun_out/ncu_q1_$TAG.log
$NCU -k regex:pipeline_kernel -s 1 -c 1 -f -o gpurun_out/prof_q1d_$TAG python profiles/run_shape.py q1d 59986052 3 > gpurun_out/ncu_q1d_$TAG.log 2>&1; tail -1 gpurun_out/ncu_q1d_$TAG.log
$NCU -k "regex:probe_pipeline|entries_pipeline" -s 4 -c 4 -f -o gpurun_out/prof_q3_sf10_$TAG python profiles/run_shape.py q3 59986052 3 > gpurun_out/ncu_q3_sf10_$TAG.log 2>&1; tail -1 gpurun_out/ncu_q3_sf10_$TAG.log
$NCU -k "regex:probe_pipeline|entries_pipeline" -s 4 -c 4 -f -o gpurun_out/prof_q3_sf100_$TAG python profiles/run_shape.py q3 600037902 3 > gpurun_out/ncu_q3_sf100_$TAG.log 2>&1; tail -1 gpurun_out/ncu_q3_sf100_$TAG.log
ls -la gpurun_out | tail -20
