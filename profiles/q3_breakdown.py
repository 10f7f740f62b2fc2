import sys, time, os
sys.path.insert(0, os.getcwd())
import torch
import pg_fusion_b200 as pg
from pg_fusion_b200 import AggFunc, Cmp, Factor
from tests import util as U
ctx = pg.Context(0)
ncust, nord, total = 15_000_000, 150_000_000, 600_037_902
cust = ctx.gen_scan(pg.GenTable.CUSTOMER_Q3, ncust, seed=42)
orders = ctx.gen_scan(pg.GenTable.ORDERS_Q3, nord, seed=42, scale_rows=ncust)
li = ctx.gen_scan(pg.GenTable.LINEITEM_Q3, total, seed=42, scale_rows=nord)
def T(label, f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize()
    print(f"{label:32s} {1e3*(time.perf_counter()-t0):8.2f} ms", getattr(r, 'kernel_ms', ''))
    return r
for rep in range(2):
    print("rep", rep)
    r1 = T("customer build", lambda: cust.pipeline().filter(1, Cmp.EQ, b"BUILDING").build_join(0, []).run())
    r2 = T("orders probe+build", lambda: orders.pipeline().filter(2, Cmp.LT, U.Q3_DATE).join(r1.join_table, 1).build_join(0, [2, 3]).run())
    p3 = (li.pipeline().filter(3, Cmp.GT, U.Q3_DATE).join(r2.join_table, 0)
          .aggregate([0, (1, 0), (1, 1)], [(AggFunc.SUM, [Factor.of(1), Factor.const_minus(1.0, 2)])], expected_groups=max(1024, r2.rows_out))
          .order_by(U.Q3_ORDER, limit=10))
    r3 = T("lineitem probe+agg+top10", lambda: p3.run())
    T("destroy t1", lambda: ctx.destroy_join_table(r1.join_table))
    T("destroy t2", lambda: ctx.destroy_join_table(r2.join_table))
