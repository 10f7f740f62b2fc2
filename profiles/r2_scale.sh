#!/bin/bash
# bench.py at N GPUs of one box (SF100, strong scaling): usage r2_scale.sh N
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 1200 $TR --master-port 29555 bench.py --gpus $N --steps 10 --warmup 3 --no-extras > gpurun_out/r2_bench_sf100_n$N.json 2> gpurun_out/r2_bench_sf100_n$N.err; echo "rc=$?"; tail -3 gpurun_out/r2_bench_sf100_n$N.err | grep -v "^\*\|OMP_NUM\|^$"; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_bench_sf100_n$N.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], {k:(round(v['ms_per_pass'],3), round(v['kernel_ms'],3)) for k,v in d['shapes'].items()}, d['parity'].get('mismatches'), 'e2e', d['e2e']['value'], d['e2e'].get('h2d_GBps_per_gpu'), d['e2e'].get('h2d_link_peak_GBps'), d['shapes']['q3'].get('nvlink_bytes_sent_per_pass_rank0'))"
