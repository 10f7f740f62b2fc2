python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python profiles/run_shape.py q1 59986052 4
