#!/bin/bash
# SF100 page-sharded strong scaling at 1/2/4/8 GPUs (run under gpurun --gpus 8)
mkdir -p gpurun_out
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then python bench.py --workload sf100 --steps 10 --warmup 3 2>gpurun_out/sf100_$n.err | tail -1 > gpurun_out/sf100_$n.json
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --workload sf100 --steps 10 --warmup 3 2>gpurun_out/sf100_$n.err | tail -1 > gpurun_out/sf100_$n.json; fi
  cat gpurun_out/sf100_$n.json | cut -c1-1500
done
