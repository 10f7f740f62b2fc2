#!/bin/bash
timeout 600 python bench.py --sf 10 --steps 3 --warmup 3 > gpurun_out/r2d_bench_sf10.json 2> gpurun_out/r2d_bench_sf10.err; echo "rc=$?"; tail -2 gpurun_out/r2d_bench_sf10.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2d_bench_sf10.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], {k:(round(v['ms_per_pass'],3), round(v['frac'],3)) for k,v in d['shapes'].items()}, 'e2e', d['e2e']['value'], d['cpu_baseline']['acero'], list(d['other_workloads'])[:3], d['parity'])"
