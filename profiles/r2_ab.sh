#!/bin/bash
for sz in 59986052 600037902; do
  for pers in "" 1; do echo "== q3 rows=$sz persist=$pers"; env ${pers:+PGF_L2_PERSIST=1} Q3_LIMIT=10 timeout 300 python profiles/run_shape.py q3 $sz 4 2>&1 | tail -1 | cut -c1-175; done
done
