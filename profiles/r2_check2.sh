#!/bin/bash
# Round 2: full parity suite on the current code, kernel timings, compute-sanitizer memcheck over the small GPU tests,
# and the SF10 bench (end-to-end figure with pipelined page hand-over).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2b_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_tests.log; tail -4 gpurun_out/r2b_tests.log
echo "== q1d"; timeout 300 python profiles/run_shape.py q1d 59986052 3 2>&1 | tail -1
echo "== q3 sf100"; Q3_LIMIT=10 timeout 300 python profiles/run_shape.py q3 600037902 4 2>&1 | tail -1
echo "== q3 sf10"; Q3_LIMIT=10 timeout 300 python profiles/run_shape.py q3 59986052 4 2>&1 | tail -1
echo "== bench sf10"; timeout 900 python bench.py --sf 10 --steps 5 --warmup 3 --no-extras --e2e-steps 3 > gpurun_out/r2b_bench_sf10.json 2> gpurun_out/r2b_bench_sf10.err; echo "rc=$?"; tail -2 gpurun_out/r2b_bench_sf10.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2b_bench_sf10.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], {k:(round(v['ms_per_pass'],3), round(v['kernel_ms'],3), round(v['frac'],3)) for k,v in d['shapes'].items()}, 'e2e', d['e2e']['value'], d['e2e'].get('h2d_GBps_per_gpu'), d['e2e'].get('h2d_link_peak_GBps'), 'cpu all cores', d['cpu_baseline']['all_cores']['value'])"
echo "== compute-sanitizer memcheck"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/r2_memcheck_raw.log python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_join.py tests/test_gpu_bloom.py tests/test_gpu_sort.py tests/test_gpu_exchange.py -m gpu -q -x -k "not sf10 and not full" > gpurun_out/r2_memcheck_pytest.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r2_memcheck_pytest.log; grep -c "Invalid\|Error" gpurun_out/r2_memcheck_raw.log; tail -5 gpurun_out/r2_memcheck_raw.log
