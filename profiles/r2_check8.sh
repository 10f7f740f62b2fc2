#!/bin/bash
timeout 900 python bench.py > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; echo "rc=$?"; tail -2 gpurun_out/bench_r2f.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_r2f.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], {k:(round(v['ms_per_pass'],3), round(v['frac'],3)) for k,v in d['shapes'].items()}, 'e2e', d['e2e']['value'], d['cpu_baseline']['acero']['q6_rows_per_s'], d['roofline']['traffic'], d['clocks']['reasons'])"
