#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q -x -k "not sf100 and not acero" 2>&1 | tail -3
echo "== q3 sf100 / sf10"; Q3_LIMIT=10 timeout 300 python profiles/run_shape.py q3 600037902 4 2>&1 | tail -1 | cut -c1-170; Q3_LIMIT=10 timeout 300 python profiles/run_shape.py q3 59986052 4 2>&1 | tail -1 | cut -c1-170
echo "== q1 generic groups"; timeout 300 python profiles/run_shape.py q1 59986052 3 2>&1 | tail -1
