#!/bin/bash
# Experiment: the orders side (join probe + build sink) split like the lineitem side (PGF_SPLIT_BUILD=1) or fused (default)
for sz in 59986052 600037902; do
  for v in "" 1; do echo "== q3 rows=$sz split build=$v"; env ${v:+PGF_SPLIT_BUILD=1} Q3_LIMIT=10 timeout 120 python profiles/run_shape.py q3 $sz 4 2>&1 | tail -1 | cut -c1-175; done
done
