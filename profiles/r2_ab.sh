#!/bin/bash
# Experiment: join tables at 2 (default), 3 (same as 4 after rounding to 2^k in most cases) and 4 slots per build row
for sz in 59986052 600037902; do
  for f in 2 4; do echo "== q3 rows=$sz capacity factor $f"; PGF_JOIN_CAPACITY_FACTOR=$f Q3_LIMIT=10 timeout 120 python profiles/run_shape.py q3 $sz 4 2>&1 | tail -1 | cut -c1-175; done
done
