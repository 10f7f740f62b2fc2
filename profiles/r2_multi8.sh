#!/bin/bash
# Round 2, N GPUs of one box: library-communicator parity, the partitioned Q3 breakdown and bench.py (SF100, strong).
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== library comm parity"; timeout 600 $TR --master-port 29541 tests/run_multi_gpu_lib.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -3
echo "== breakdown sf100"; timeout 600 $TR --master-port 29552 profiles/q3_partitioned_breakdown.py 100 2>&1 | grep -v "^\*\|OMP_NUM\|^$\|NCCL version" | tail -14
echo "== bench sf100 N=$N"; timeout 1200 $TR --master-port 29553 bench.py --gpus $N --steps 10 --warmup 3 --no-extras > gpurun_out/r2_bench_sf100_n$N.json 2> gpurun_out/r2_bench_sf100_n$N.err; echo "rc=$?"; tail -3 gpurun_out/r2_bench_sf100_n$N.err | grep -v "^\*\|OMP_NUM\|^$"; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_bench_sf100_n$N.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], {k:(round(v['ms_per_pass'],3), round(v['kernel_ms'],3)) for k,v in d['shapes'].items()}, d['parity'].get('mismatches'), 'e2e', d['e2e']['value'], d['e2e'].get('h2d_GBps_per_gpu'), d['e2e'].get('h2d_link_peak_GBps'), d['shapes']['q3'].get('nvlink_bytes_sent_per_pass_rank0'), d['wall_ms_per_step'])"
