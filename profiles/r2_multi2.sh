#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== pytest (top-k, joins, exchange)"; timeout 900 python -m pytest tests/test_gpu_sort.py tests/test_gpu_join.py tests/test_gpu_exchange.py tests/test_gpu_multi_rank.py -m gpu -q 2>&1 | tail -3
echo "== library comm parity"; timeout 900 $TR --master-port 29541 tests/run_multi_gpu_lib.py 2>&1 | tail -1
port=29550
for sf in 10 100; do port=$((port+1)); echo "== breakdown sf$sf"; timeout 600 $TR --master-port $port profiles/q3_partitioned_breakdown.py $sf 2>&1 | grep -v "^\*\|OMP_NUM\|^$\|NCCL version" | tail -14; done
for sf in 10 100; do port=$((port+1))
echo "== bench sf$sf N=$N"; timeout 1200 $TR --master-port $port bench.py --gpus $N --sf $sf --steps 5 --warmup 3 --no-extras > gpurun_out/r2_bench_sf${sf}_n$N.json 2> gpurun_out/r2_bench_sf${sf}_n$N.err; tail -3 gpurun_out/r2_bench_sf${sf}_n$N.err | grep -v "^\*\|OMP_NUM\|^$"; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_bench_sf${sf}_n$N.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], {k:v['ms_per_pass'] for k,v in d['shapes'].items()}, d['parity'].get('mismatches'), d['e2e']['value'], d['shapes']['q3'].get('nvlink_bytes_sent_per_pass_rank0'))"
done
