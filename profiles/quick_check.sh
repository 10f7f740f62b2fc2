python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python profiles/q3_breakdown.py 2>&1 | tail -6
python bench.py --workload sf100 --steps 5 --warmup 2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['q3_no_bloom']); print(d['q3_bloom_16_bits_per_key'])"
