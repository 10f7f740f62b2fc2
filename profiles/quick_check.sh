python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python profiles/run_shape.py q3 59986052 3 | tail -1
python profiles/run_shape.py q1 59986052 3 | tail -1
python profiles/run_shape.py q1d 59986052 3 | tail -1
