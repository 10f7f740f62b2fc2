#!/usr/bin/env python
"""Run one plan shape a few times (for ncu captures): python profiles/run_shape.py {q6|q1} [rows] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pg_fusion_b200 as pg
from tests import util as U

shape = sys.argv[1] if len(sys.argv) > 1 else "q6"
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 12_000_000
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 4
with pg.Context(0) as ctx:
    if shape == "q6":
        scan = ctx.gen_scan(pg.GenTable.LINEITEM_Q6, rows, seed=42); plan = U.gpu_q6(scan); bpr = 40
    else:
        scan = ctx.gen_scan(pg.GenTable.LINEITEM_Q1, rows, seed=42); plan = U.gpu_q1(scan); bpr = 80
    for i in range(iters):
        r = plan.run()
        print(f"{shape} iter {i}: kernel {r.kernel_ms:.4f} ms, {rows * bpr / r.kernel_ms / 1e6:.1f} GB/s, rows_out {r.rows_out}")
