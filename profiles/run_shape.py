#!/usr/bin/env python
"""Run one plan shape a few times (for ncu captures):
    python profiles/run_shape.py {q6|q1|q3|q3bloom|bloom} [rows] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pg_fusion_b200 as pg
from tests import util as U

shape = sys.argv[1] if len(sys.argv) > 1 else "q6"
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 12_000_000
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 4


def pow2(n):
    b = 1
    while b < n:
        b <<= 1
    return b


with pg.Context(0) as ctx:
    if shape in ("q6d", "q1d"):
        if shape == "q6d":
            scan = ctx.gen_scan(pg.GenTable.LINEITEM_Q6_D, rows, seed=42); plan = U.gpu_q6_d(scan); bpr = 52
        else:
            scan = ctx.gen_scan(pg.GenTable.LINEITEM_Q1_D, rows, seed=42); plan = U.gpu_q1_d(scan); bpr = 72
        for i in range(iters):
            r = plan.run()
            print(f"{shape} iter {i}: kernel {r.kernel_ms:.4f} ms, {rows * bpr / r.kernel_ms / 1e6:.1f} GB/s, rows_out {r.rows_out}")
    elif shape in ("q6", "q1"):
        if shape == "q6":
            scan = ctx.gen_scan(pg.GenTable.LINEITEM_Q6, rows, seed=42); plan = U.gpu_q6(scan); bpr = 40
        else:
            scan = ctx.gen_scan(pg.GenTable.LINEITEM_Q1, rows, seed=42); plan = U.gpu_q1(scan); bpr = 80
        for i in range(iters):
            r = plan.run()
            print(f"{shape} iter {i}: kernel {r.kernel_ms:.4f} ms, {rows * bpr / r.kernel_ms / 1e6:.1f} GB/s, rows_out {r.rows_out}")
    elif shape in ("q3", "q3bloom"):
        scale = rows / 59_986_052
        ncust, nord = max(1000, int(1_500_000 * scale)), max(10000, int(15_000_000 * scale))
        cust = ctx.gen_scan(pg.GenTable.CUSTOMER_Q3, ncust, seed=42)
        orders = ctx.gen_scan(pg.GenTable.ORDERS_Q3, nord, seed=42, scale_rows=ncust)
        li = ctx.gen_scan(pg.GenTable.LINEITEM_Q3, rows, seed=42, scale_rows=nord)
        bp = None
        if shape == "q3bloom":
            bp = (pg.BloomParams.new(pow2(16 * ncust // 5), 4, 7), pg.BloomParams.new(pow2(16 * nord // 10), 4, 7))
        import time
        for i in range(iters):
            t0 = time.perf_counter()
            res, st = U.gpu_q3(ctx, cust, orders, li, bp, limit=int(os.environ.get("Q3_LIMIT", "0")))
            wall = (time.perf_counter() - t0) * 1e3
            t = (st["customer"].kernel_ms, st["orders"].kernel_ms, st["lineitem"].kernel_ms)
            print(f"  whole pass {wall:.3f} ms (host clock, limit={os.environ.get('Q3_LIMIT', '0')}); ", end="")
            print(f"{shape} iter {i}: kernels {t[0]:.4f} {t[1]:.4f} {t[2]:.4f} ms; lineitem {rows * 36 / t[2] / 1e6:.1f} GB/s; "
                  f"bloom->{st['lineitem'].rows_bloom} filter->{st['lineitem'].rows_filtered} joined {st['lineitem'].rows_out} groups {len(res.keys)}")
    elif shape == "bloom":
        p = pg.BloomParams.new(**pg.GUC_DEFAULT_BLOOM)
        keys = ctx.gen_scan(pg.GenTable.KEYS_I64, 1_000_000, seed=7)
        probe = ctx.gen_scan(pg.GenTable.KEYS_I64, rows, seed=7)
        rf = ctx.runtime_filter(p)
        rf.try_acquire_builder(); rf.insert_scan(keys, 0)
        print(f"bloom build: {ctx.last_kernel_ms():.4f} ms")
        rf.publish_ready()
        for i in range(iters):
            d, st = rf.probe_scan(probe, 0)
            k = ctx.last_kernel_ms()
            print(f"bloom probe iter {i}: kernel {k:.4f} ms, {rows / k / 1e6:.2f} G probes/s, {rows * 9 / k / 1e6:.1f} GB/s, rejected {st.rejected_rows}")
    elif shape == "q3var":
        # isolate the cost components of the Q3 lineitem pipeline (probe side)
        from pg_fusion_b200 import AggFunc, Cmp, Factor
        scale = rows / 59_986_052
        ncust, nord = max(1000, int(1_500_000 * scale)), max(10000, int(15_000_000 * scale))
        cust = ctx.gen_scan(pg.GenTable.CUSTOMER_Q3, ncust, seed=42)
        orders = ctx.gen_scan(pg.GenTable.ORDERS_Q3, nord, seed=42, scale_rows=ncust)
        li = ctx.gen_scan(pg.GenTable.LINEITEM_Q3, rows, seed=42, scale_rows=nord)
        r1 = cust.pipeline().filter(1, Cmp.EQ, b"BUILDING").build_join(0, []).run()
        full = orders.pipeline().filter(2, Cmp.LT, U.Q3_DATE).join(r1.join_table, 1).build_join(0, [2, 3]).run()
        empty = orders.pipeline().filter(2, Cmp.LT, b"0000-00-00").build_join(0, [2, 3]).run()
        allord = orders.pipeline().build_join(0, [2, 3]).run()
        for label, table, date in (("full", full, U.Q3_DATE), ("nothing passes the filter", full, b"9999-99-99"),
                                   ("empty build side", empty, U.Q3_DATE), ("all orders build side (every probe matches)", allord, U.Q3_DATE)):
            p3 = (li.pipeline().filter(3, Cmp.GT, date).join(table.join_table, 0)
                  .aggregate([0, (1, 0), (1, 1)], [(AggFunc.SUM, [Factor.of(1), Factor.const_minus(1.0, 2)])],
                             expected_groups=max(1024, table.rows_out)))
            for i in range(iters):
                r = p3.run()
            print(f"q3var [{label}]: kernel {r.kernel_ms:.4f} ms, filter->{r.rows_filtered} joined {r.rows_out} groups {len(r.keys)}")
