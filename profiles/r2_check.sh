#!/bin/bash
# Round 2: parity suite + Q3 kernel times + the bench end to end (1 GPU).
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2_tests6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests6.log
tail -8 gpurun_out/r2_tests6.log
for sz in 59986052 600037902; do echo "== q3 rows=$sz"; timeout 300 python profiles/run_shape.py q3 $sz 3 2>&1 | tail -1; PGF_PROBE_PREFETCH=0 timeout 300 python profiles/run_shape.py q3 $sz 3 2>&1 | tail -1; done
echo "== q3bloom"; timeout 300 python profiles/run_shape.py q3bloom 600037902 3 2>&1 | tail -1
echo "== q6 / q1 sf10"; timeout 300 python profiles/run_shape.py q6 59986052 3 2>&1 | tail -1; timeout 300 python profiles/run_shape.py q1 59986052 3 2>&1 | tail -1
echo "== q1 sf100"; timeout 300 python profiles/run_shape.py q1 600037902 3 2>&1 | tail -1
echo "== bench ours"; timeout 1500 python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/r2_bench_try.json 2> gpurun_out/r2_bench_try.err; tail -5 gpurun_out/r2_bench_try.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_bench_try.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], {k:(round(v['ms_per_pass'],3), round(v['kernel_ms'],3), round(v['frac'],3)) for k,v in d['shapes'].items()}, d['parity'].get('mismatches'), d['e2e']['value'], d['shapes']['q3']['kernel_ms_by_pipeline'], d['shapes']['q3']['ms_per_pass_with_runtime_filters'])"
