"""Multi-GPU parity run (one process per GPU, NCCL): python -m torch.distributed.run --nproc-per-node N
--master-addr 127.0.0.1 tests/run_multi_gpu.py.  Every rank holds a page shard of each scan; the
sharded Q6 / Q1 / Q3 results must equal the single-GPU results of the library and the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pg_fusion_b200 as pg  # noqa: E402
from oracle import pyorc as O  # noqa: E402
from pg_fusion_b200 import GenTable  # noqa: E402
from pg_fusion_b200 import multi_gpu as MG  # noqa: E402
from tests import util as U  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    ctx = pg.Context(local)
    # ---- Q6 / Q1: page shards + all-gather of partial states + rank-order merge
    rows = 200_000   # Float64 sums stay within 1e-12 of the sequential (reference-order) oracle at this size
    for table, schema, gpu, orc in ((GenTable.LINEITEM_Q6, U.Q6_SCHEMA, U.gpu_q6, U.oracle_q6),
                                    (GenTable.LINEITEM_Q1, U.Q1_SCHEMA, U.gpu_q1, U.oracle_q1)):
        lo, hi = MG.shard_range(rows, rank, world)
        shard = ctx.gen_scan(table, hi - lo, seed=42, first_row=lo)
        merged, stats = MG.merge_partial_aggregate(gpu(shard), world, device, max_groups=64)
        assert stats.rows_in == hi - lo
        whole = ctx.gen_scan(table, rows, seed=42)
        single = gpu(whole).run()
        U.assert_agg_equal(merged, single)
        if rank == 0:
            want = orc(O.OTable.from_pages(whole.read_pages(), 65536, U.orc_cols(schema)))
            U.assert_agg_equal(merged, want)
        shard.release()
        whole.release()
    # ---- Q3: broadcast joins, OR-merged Bloom filters, partial/final GROUP BY
    ncust, nord, nli = 15_000, 150_000, 600_000
    def pow2(n):
        b = 1
        while b < n:
            b <<= 1
        return b
    for bp in (None, (pg.BloomParams.new(pow2(16 * ncust // 5), 4, 7), pg.BloomParams.new(pow2(16 * nord // 10), 4, 7))):
        shards, wholes = [], []
        for table, n, scale in ((GenTable.CUSTOMER_Q3, ncust, 0), (GenTable.ORDERS_Q3, nord, ncust), (GenTable.LINEITEM_Q3, nli, nord)):
            lo, hi = MG.shard_range(n, rank, world)
            shards.append(ctx.gen_scan(table, hi - lo, seed=42, first_row=lo, scale_rows=scale))
            wholes.append(ctx.gen_scan(table, n, seed=42, scale_rows=scale))
        res, st = U.gpu_q3_sharded(ctx, *shards, world, device, bp)
        single, st1 = U.gpu_q3(ctx, *wholes, bp)
        assert len(res.keys) == len(single.keys) > 0
        U.assert_agg_equal(res, single)
        joined = MG.all_gather_counts(st["lineitem"].rows_out, world, device)
        assert sum(joined) == st1["lineitem"].rows_out, (joined, st1["lineitem"].rows_out)
        assert U.top10(res)[0][0] == U.top10(single)[0][0]
        # ORDER BY revenue DESC, o_orderdate LIMIT 10 on the merged groups (device top-k after the merge)
        top, _ = U.gpu_q3_sharded(ctx, *shards, world, device, bp, limit=10)
        want10 = U.top10(single)
        assert [k[0] for k in top.keys] == [r[0] for r in want10]
        for a, w in zip(top.aggs, want10):
            U.assert_close(a[0], w[1], 1e-12, "revenue")
        for s in shards + wholes:
            s.release()
    dist.barrier()
    if rank == 0:
        print(f"multi-GPU parity ok on {world} ranks")
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
