python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python profiles/run_shape.py q3 59986052 3 | tail -1
python profiles/run_shape.py q3bloom 59986052 3 | tail -1
