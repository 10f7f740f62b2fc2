python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python profiles/run_shape.py q1 59986052 4
python profiles/run_shape.py bloom 64000000 4
python profiles/run_shape.py q3bloom 59986052 3
python profiles/run_shape.py q6 59986052 3
