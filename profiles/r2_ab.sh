#!/bin/bash
# Experiment: the Q3 lineitem pipeline with a group table sized by the true group count instead of the orders build side
for hint in 0 1500000 3000000; do echo "== sf100 groups hint $hint"; Q3_GROUPS_HINT=$hint Q3_LIMIT=10 timeout 300 python profiles/run_shape.py q3 600037902 4 2>&1 | tail -1 | cut -c1-160; done
for hint in 0 150000; do echo "== sf10 groups hint $hint"; Q3_GROUPS_HINT=$hint Q3_LIMIT=10 timeout 300 python profiles/run_shape.py q3 59986052 4 2>&1 | tail -1 | cut -c1-160; done
