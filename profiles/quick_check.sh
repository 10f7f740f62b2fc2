python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python profiles/run_shape.py q3var 59986052 3
python profiles/run_shape.py q3bloom 59986052 3
