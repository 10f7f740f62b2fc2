#!/bin/bash
# Round 2, N GPUs of one box: library-communicator parity (tests/run_multi_gpu_lib.py), the round-1 host-driven path,
# the plain-C sharded example, and bench.py at N = 1 and N (SF10 quick, then SF100).  usage: r2_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== library comm parity"; timeout 900 $TR --master-port 29541 tests/run_multi_gpu_lib.py 2>&1 | tail -5
echo "== host-driven parity (round 1 path)"; timeout 900 $TR --master-port 29542 tests/run_multi_gpu.py 2>&1 | tail -3
echo "== pytest multi"; timeout 900 python -m pytest tests/test_gpu_multi_rank.py tests/test_c_abi_example.py -m gpu -q 2>&1 | tail -4
echo "== bench sf10 N=1"; timeout 900 python bench.py --sf 10 --steps 5 --warmup 3 --no-extras --record-expected > gpurun_out/r2_bench_sf10_n1.json 2> gpurun_out/r2_bench_sf10_n1.err; tail -3 gpurun_out/r2_bench_sf10_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_sf10_n1.json')); print(d['value'], d['ms_per_step'], {k:v['ms_per_pass'] for k,v in d['shapes'].items()}, d['e2e']['value'])"
cp profiles/sf_expected.json gpurun_out/sf_expected.json
echo "== bench sf10 N=$N"; timeout 900 $TR --master-port 29543 bench.py --gpus $N --sf 10 --steps 5 --warmup 3 --no-extras > gpurun_out/r2_bench_sf10_n$N.json 2> gpurun_out/r2_bench_sf10_n$N.err; tail -3 gpurun_out/r2_bench_sf10_n$N.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_sf10_n$N.json')); print(d['value'], d['ms_per_step'], {k:v['ms_per_pass'] for k,v in d['shapes'].items()}, d['parity'], d['e2e']['value'], d['shapes']['q3'].get('nvlink_bytes_sent_per_pass_rank0'))"
echo "== bench sf100 N=$N"; timeout 1200 $TR --master-port 29544 bench.py --gpus $N --steps 5 --warmup 3 --no-extras > gpurun_out/r2_bench_sf100_n$N.json 2> gpurun_out/r2_bench_sf100_n$N.err; tail -3 gpurun_out/r2_bench_sf100_n$N.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_sf100_n$N.json')); print(d['value'], d['ms_per_step'], {k:v['ms_per_pass'] for k,v in d['shapes'].items()}, d['parity'], d['e2e']['value'], d['shapes']['q3'].get('nvlink_bytes_sent_per_pass_rank0'))"
