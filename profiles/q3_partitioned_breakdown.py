#!/usr/bin/env python
"""Where the time of the hash-partitioned Q3 plan goes (torchrun, N ranks):
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/q3_partitioned_breakdown.py [sf]
Wall-clock per stage on rank 0 with a device synchronisation after each (so the sum exceeds the pipelined total)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist
import pg_fusion_b200 as pg
from pg_fusion_b200 import AggFunc, Cmp, ColumnSpec, Factor, GenTable, TypeTag
from pg_fusion_b200 import multi_gpu as MG, tpch as T

sf = int(sys.argv[1]) if len(sys.argv) > 1 else 10
nli, nord, ncust = {10: (59_986_052, 15_000_000, 1_500_000), 100: (600_037_902, 150_000_000, 15_000_000)}[sf]
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("gloo")
ctx = pg.Context(local)
ids = [pg.Context.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
ctx.comm_init(ids[0], rank, world)
sh = lambda n: MG.shard_range(n, rank, world)
scans = []
for table, n, scale in ((GenTable.CUSTOMER_Q3, ncust, 0), (GenTable.ORDERS_Q3, nord, ncust), (GenTable.LINEITEM_Q3, nli, nord)):
    lo, hi = sh(n)
    scans.append(ctx.gen_scan(table, hi - lo, seed=42, first_row=lo, scale_rows=scale))
cust, orders, li = scans
rf = ctx.runtime_filter(T.q3_bloom_params(1, nord)[1])
for _ in range(3):
    top, st = T.gpu_q3_partitioned(ctx, cust, orders, li, nord_total=nord, limit=10, rf=rf)
ctx.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(5):
    top, st = T.gpu_q3_partitioned(ctx, cust, orders, li, nord_total=nord, limit=10, rf=rf)
ctx.synchronize(); dist.barrier()
total = (time.perf_counter() - t0) / 5
marks = []
def mark(name, t=[None]):
    ctx.synchronize()
    now = time.perf_counter()
    if t[0] is not None:
        marks.append((name, (now - t[0]) * 1e3))
    t[0] = now
hint = lambda scan, frac: int(scan.info().rows * frac) + 4096
mark("start")
r1 = cust.pipeline().filter(1, Cmp.EQ, b"BUILDING").build_join(0, [], rows_only=True, expected_rows=hint(cust, 0.25)).run(); mark("customer rows")
t1, _ = ctx.exchange(r1.join_table, partition=False); mark("customer broadcast + table")
rf.retire_ready_after_quiescence(); rf.try_acquire_builder(); mark("filter clear")
r2 = orders.pipeline().filter(2, Cmp.LT, T.Q3_DATE).join(t1, 1).build_join(0, [2, 3], rf, rows_only=True, expected_rows=hint(orders, 0.125)).run(); mark("orders rows (+ filter build)")
t2, sent2 = ctx.exchange(r2.join_table, partition=True); mark("orders partition + table")
rf.or_all_reduce(); rf.publish_ready(); mark("filter OR all-reduce + publish")
r3 = li.pipeline().bloom_probe(rf, 0).filter(3, Cmp.GT, T.Q3_DATE).build_join(0, [1, 2], rows_only=True, expected_rows=hint(li, 1 / 30)).run(); mark("lineitem filter + Bloom -> rows")
rs3, sent3 = ctx.exchange(r3.join_table, partition=True, rows_only=True); mark("lineitem rows partition")
schema = [ColumnSpec(TypeTag.Int32), ColumnSpec(TypeTag.Float64), ColumnSpec(TypeTag.Float64)]
r4 = (ctx.row_set_pipeline(rs3, schema).join(t2, 0).aggregate([0, (1, 0), (1, 1)], [(AggFunc.SUM, [Factor.of(1), Factor.const_minus(1.0, 2)])],
      expected_groups=max(1024, min(ctx.join_table_info(t2).rows, ctx.join_table_info(rs3).rows))).order_by(T.Q3_ORDER, limit=10).run()); mark("probe + GROUP BY + top-10")
if rank == 0:
    print(f"SF{sf}, {world} ranks: partitioned Q3 = {total * 1e3:.3f} ms per pass (pipelined)")
    for name, ms in marks:
        print(f"  {name:36s} {ms:8.3f} ms")
    print(f"  kernels: customer {r1.kernel_ms:.3f}  orders {r2.kernel_ms:.3f}  lineitem {r3.kernel_ms:.3f}  final {r4.kernel_ms:.3f} ms; "
          f"lineitem rows routed {r3.rows_out} of {r3.rows_in} (Bloom rejected {r3.rows_in - r3.rows_bloom}); NVLink bytes sent: orders {sent2}, lineitem {sent3}")
ctx.comm_destroy(); ctx.close(); dist.destroy_process_group()
