"""Multi-GPU parity run with every collective INSIDE the library (pgf_comm_*: NCCL over NVLink behind the C ABI):
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/run_multi_gpu_lib.py
torch.distributed (gloo) is used for one thing only: handing rank 0's communicator id to the other ranks.
Checked: sharded Q6 / Q1 (pgf_pipeline_run_sharded) against the single-GPU run and the oracle, the Bloom OR
all-reduce bit for bit, broadcast and hash-partitioned join exchanges, and the hash-partitioned Q3 plan
 against the single-GPU Q3 and the oracle."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pg_fusion_b200 as pg  # noqa: E402
from oracle import pyorc as O  # noqa: E402
from pg_fusion_b200 import AggFunc, Cmp, ColumnSpec, Factor, GenTable, TypeTag, _lib  # noqa: E402
from pg_fusion_b200 import multi_gpu as MG  # noqa: E402
from pg_fusion_b200 import tpch as T  # noqa: E402
from tests import util as U  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("gloo")
    ctx = pg.Context(local)
    ids = [pg.Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.comm_init(ids[0], rank, world)
    assert ctx.comm_info() == (rank, world)

    # ---- Q6 / Q1: Partial -> all-gather -> Final in one library call
    rows = 200_000
    for table, schema, plan, orc, groups in ((GenTable.LINEITEM_Q6, U.Q6_SCHEMA, T.gpu_q6, U.oracle_q6, 1),
                                             (GenTable.LINEITEM_Q1, U.Q1_SCHEMA, T.gpu_q1, U.oracle_q1, 16)):
        lo, hi = MG.shard_range(rows, rank, world)
        shard = ctx.gen_scan(table, hi - lo, seed=42, first_row=lo)
        merged = plan(shard).run_sharded(max_groups=groups)
        whole = ctx.gen_scan(table, rows, seed=42)
        U.assert_agg_equal(merged, plan(whole).run())
        U.assert_agg_equal(merged, orc(O.OTable.from_pages(whole.read_pages(), 65536, U.orc_cols(schema))))
        again = plan(shard).run_sharded(max_groups=groups)
        assert again.by_key() == merged.by_key()          # fixed-order merge: the same bits every time, on every rank
        shard.release()
        whole.release()

    # ---- Bloom OR all-reduce: every rank inserts its shard of the keys; the merged filter is the filter of all keys.
    # 2^18 bits take the all-gather path, 2^24 bits (2 MiB) the reduce-scatter + all-gather path of large filters.
    for bits, nkeys in ((1 << 18, 100_000), (1 << 24, 1_000_000)):
        p = pg.BloomParams.new(bits, 4, 7)
        lo, hi = MG.shard_range(nkeys, rank, world)
        keys = ctx.gen_scan(GenTable.KEYS_I64, hi - lo, seed=7, first_row=lo)
        allkeys = ctx.gen_scan(GenTable.KEYS_I64, nkeys, seed=7)
        rf, ref = ctx.runtime_filter(p), ctx.runtime_filter(p)
        rf.try_acquire_builder(); ref.try_acquire_builder()
        rf.insert_scan(keys, 0); ref.insert_scan(allkeys, 0)
        rf.or_all_reduce()
        rf.publish_ready(); ref.publish_ready()
        assert (rf.words() == ref.words()).all(), f"Bloom OR all-reduce of {bits} bits"
        keys.release(); allkeys.release()

    # ---- join exchanges on a plain key / payload table
    n = 50_000
    lo, hi = MG.shard_range(n, rank, world)
    mine = ctx.gen_scan(GenTable.ORDERS_Q3, hi - lo, seed=42, first_row=lo, scale_rows=1000)
    everything = ctx.gen_scan(GenTable.ORDERS_Q3, n, seed=42, scale_rows=1000)
    b = mine.pipeline().build_join(0, [1, 3], rows_only=True).run()
    assert b.rows_out == hi - lo
    full, sent_b = ctx.exchange(b.join_table, partition=False)
    assert ctx.join_table_info(full).rows == n
    part, sent_p = ctx.exchange(b.join_table, partition=True)
    owned = ctx.join_table_info(part).rows
    counts = [int.from_bytes(x, "little") for x in ctx.comm_all_gather_host(owned.to_bytes(8, "little"))]
    assert sum(counts) == n and all(abs(c - n / world) < 0.1 * n / world + 50 for c in counts), counts
    # every key of the whole table finds its row in the broadcast table; in the partitioned one exactly the keys this rank owns
    probe_all = everything.pipeline().join(full, 0).aggregate([], [(AggFunc.COUNT_STAR, None), (AggFunc.SUM, [Factor.of((1, 0))])]).run()
    want = everything.pipeline().aggregate([], [(AggFunc.COUNT_STAR, None), (AggFunc.SUM, [Factor.of(1)])]).run()
    assert probe_all.aggs[0] == want.aggs[0]
    probe_part = everything.pipeline().join(part, 0).aggregate([0], [(AggFunc.COUNT_STAR, None)], expected_groups=n).run()
    L = _lib.lib()
    assert len(probe_part.keys) == owned and all(L.pgf_partition_of_key(k[0], world) == rank for k in probe_part.keys)
    for h in (b.join_table, full, part):
        ctx.destroy_join_table(h)
    mine.release(); everything.release()

    # ---- the hash-partitioned Q3 plan
    ncust, nord, nli = 15_000, 150_000, 600_000
    shards, wholes = [], []
    for table, total, scale in ((GenTable.CUSTOMER_Q3, ncust, 0), (GenTable.ORDERS_Q3, nord, ncust), (GenTable.LINEITEM_Q3, nli, nord)):
        lo, hi = MG.shard_range(total, rank, world)
        shards.append(ctx.gen_scan(table, hi - lo, seed=42, first_row=lo, scale_rows=scale))
        wholes.append(ctx.gen_scan(table, total, seed=42, scale_rows=scale))
    single, st1 = T.gpu_q3(ctx, *wholes)
    want10 = U.top10(single)
    top, st = T.gpu_q3_partitioned(ctx, *shards, nord_total=nord, limit=10)
    assert [(r[0], r[2], r[3]) for r in top] == [(r[0], r[2], r[3]) for r in want10], (top[:3], want10[:3])
    for g, w in zip(top, want10):
        U.assert_close(g[1], w[1], 1e-12, "revenue")
    # without the LIMIT: this rank's groups are exactly the groups of the keys it owns, with the single-GPU sums
    mine_rows, st = T.gpu_q3_partitioned(ctx, *shards, nord_total=nord, limit=0)
    ref = single.by_key()
    owned_ref = {k: v for k, v in ref.items() if L.pgf_partition_of_key(k[0], world) == rank}
    assert len(mine_rows) == len(owned_ref)
    for k, v, d, pr in mine_rows:
        U.assert_close(v, owned_ref[(k, d, pr)][0], 1e-12, f"group {k}")
    joined = [int.from_bytes(x, "little") for x in ctx.comm_all_gather_host(int(st["final"].rows_out).to_bytes(8, "little"))]
    assert sum(joined) == st1["lineitem"].rows_out
    crossed = [int.from_bytes(x, "little") for x in ctx.comm_all_gather_host(int(st["lineitem"].rows_out).to_bytes(8, "little"))]
    assert sum(crossed) < 0.05 * nli          # only what the runtime filter cannot rule out crosses NVLink
    if rank == 0:
        tables = [O.OTable.from_pages(s.read_pages(), 65536, U.orc_cols(sc)) for s, sc in
                  zip(wholes, (U.CUSTOMER_SCHEMA, U.ORDERS_SCHEMA, U.LINEITEM_Q3_SCHEMA))]
        orc, _ = U.oracle_q3(*tables)
        o10 = U.top10(orc)
        assert [(r[0], r[2], r[3]) for r in top] == [(r[0], r[2], r[3]) for r in o10]
    for s in shards + wholes:
        s.release()
    dist.barrier()
    if rank == 0:
        print(f"library-comm multi-GPU parity ok on {world} ranks (rows crossing NVLink for Q3: {sum(crossed)} of {nli})")
    ctx.comm_destroy()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
