#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q -x -k "not sf100 and not acero" 2>&1 | tail -3
echo "== phase trace q3 sf100"; Q3_LIMIT=10 PGF_TRACE=1 timeout 300 python profiles/run_shape.py q3 600037902 2 2>&1 | tail -5
echo "== q3 sf100 / sf10"; Q3_LIMIT=10 timeout 300 python profiles/run_shape.py q3 600037902 4 2>&1 | tail -1; Q3_LIMIT=10 timeout 300 python profiles/run_shape.py q3 59986052 4 2>&1 | tail -1
