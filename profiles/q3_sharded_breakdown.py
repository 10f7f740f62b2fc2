"""Wall-clock breakdown of the sharded Q3 query at SF100 (torchrun --nproc-per-node N profiles/q3_sharded_breakdown.py)."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch, torch.distributed as dist
import pg_fusion_b200 as pg
from pg_fusion_b200 import AggFunc, Cmp, Factor, multi_gpu as MG
from tests import util as U
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); device = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=device)
ctx = pg.Context(local)
ncust, nord, total = 15_000_000, 150_000_000, 600_037_902
shards = []
for table, n, scale in ((pg.GenTable.CUSTOMER_Q3, ncust, 0), (pg.GenTable.ORDERS_Q3, nord, ncust), (pg.GenTable.LINEITEM_Q3, total, nord)):
    a, b = MG.shard_range(n, rank, world)
    shards.append(ctx.gen_scan(table, b - a, seed=42, first_row=a, scale_rows=scale))
cust, orders, li = shards
def T(label, f):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); dist.barrier()
    if rank == 0: print(f"{label:36s} {1e3*(time.perf_counter()-t0):8.2f} ms")
    return r
for rep in range(2):
    if rank == 0: print("rep", rep)
    r1 = T("customer build (local)", lambda: cust.pipeline().filter(1, Cmp.EQ, b"BUILDING").build_join(0, []).run())
    t1 = T("customer table broadcast", lambda: MG.broadcast_join_table(ctx, r1.join_table, world, device))
    r2 = T("orders probe+build (local)", lambda: orders.pipeline().filter(2, Cmp.LT, U.Q3_DATE).join(t1, 1).build_join(0, [2, 3]).run())
    t2 = T("orders table broadcast", lambda: MG.broadcast_join_table(ctx, r2.join_table, world, device))
    tot = ctx.join_table_info(t2).rows
    p3 = (li.pipeline().filter(3, Cmp.GT, U.Q3_DATE).join(t2, 0)
          .aggregate([0, (1, 0), (1, 1)], [(AggFunc.SUM, [Factor.of(1), Factor.const_minus(1.0, 2)])], expected_groups=max(1024, tot))
          .order_by(U.Q3_ORDER, limit=10))
    res = T("lineitem partial + gather + merge", lambda: MG.merge_partial_aggregate(p3, world, device, max_groups=max(1024, tot)))
    T("destroy tables", lambda: (ctx.destroy_join_table(t1), ctx.destroy_join_table(t2)))
ctx.close(); dist.destroy_process_group()
