set -x
time python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
time python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
time python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-400
time python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.json
