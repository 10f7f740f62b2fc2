for v in libv6 libv7; do echo "== $v"; PGF_B200_LIB=$PWD/pg_fusion_b200/$v.so python profiles/run_shape.py q3var 59986052 3 2>&1 | grep -v "all orders"; done
