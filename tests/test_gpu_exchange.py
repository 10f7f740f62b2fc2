"""SURVEY 8e rows 4-5 on ONE GPU: row sets (PGF_BUILD_ROWS_ONLY), row-set scans and the hash-partitioned exchange
(PGF_XCHG_PARTITION), with the ranks of the partition played one after the other by a context without a
communicator (PGF_XCHG_EMULATE: the same count / scatter kernels a real exchange runs before its all-to-all).
The N-rank NCCL run of the same plan is tests/run_multi_gpu_lib.py (needs >= 2 GPUs) and bench.py --gpus N."""
import numpy as np
import pytest

import pg_fusion_b200 as pg
from oracle import pyorc as O
from pg_fusion_b200 import AggFunc, Cmp, ColumnSpec, Factor, GenTable, TypeTag, _lib
from pg_fusion_b200 import tpch as T

from . import util as U

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = pg.Context()
    yield c
    c.close()


def test_row_set_scan_equals_the_page_scan(ctx):
    """orders -> row set {o_orderkey; o_custkey, o_shippriority}; a pipeline over the row set computes what the same
    pipeline over the pages computes (keys, payload columns, a join probe on a payload column, exact Int64 sums)."""
    n, ncust = 200_000, 20_000
    orders = ctx.gen_scan(GenTable.ORDERS_Q3, n, seed=42, scale_rows=ncust)
    cust = ctx.gen_scan(GenTable.CUSTOMER_Q3, ncust, seed=42)
    t1 = cust.pipeline().filter(1, Cmp.EQ, b"BUILDING").build_join(0, []).run()
    rows = orders.pipeline().filter(2, Cmp.LT, U.Q3_DATE).build_join(0, [1, 3], rows_only=True).run()
    info = ctx.join_table_info(rows.join_table)
    assert info.rows == rows.rows_out == rows.rows_filtered and info.capacity == 0
    schema = [ColumnSpec(TypeTag.Int32)] * 3
    via_rows = (ctx.row_set_pipeline(rows.join_table, schema).join(t1.join_table, 1)
                .aggregate([2], [(AggFunc.COUNT_STAR, None), (AggFunc.SUM, [Factor.of(0)]), (AggFunc.SUM, [Factor.of(1)])]).run())
    via_pages = (orders.pipeline().filter(2, Cmp.LT, U.Q3_DATE).join(t1.join_table, 1)
                 .aggregate([3], [(AggFunc.COUNT_STAR, None), (AggFunc.SUM, [Factor.of(0)]), (AggFunc.SUM, [Factor.of(1)])]).run())
    assert via_rows.variant == "row_set_scan" and via_rows.rows_in == info.rows
    assert via_rows.by_key() == via_pages.by_key() and via_rows.rows_out == via_pages.rows_out > 0
    # a row set has no hash table: it cannot be probed
    with pytest.raises(pg.PgfError):
        orders.pipeline().join(rows.join_table, 0).count().run()
    ctx.destroy_join_table(rows.join_table)
    ctx.destroy_join_table(t1.join_table)
    orders.release(); cust.release()


@pytest.mark.parametrize("world", [2, 8])
def test_partition_is_a_disjoint_cover_by_owner(ctx, world):
    n = 300_000
    orders = ctx.gen_scan(GenTable.ORDERS_Q3, n, seed=42, scale_rows=1000)
    rows = orders.pipeline().build_join(0, [3], rows_only=True).run()
    L = _lib.lib()
    seen, total = set(), 0
    for r in range(world):
        part, _ = ctx.exchange(rows.join_table, partition=True, emulate=(world, r))
        got = orders.pipeline().join(part, 0).aggregate([0], [(AggFunc.COUNT_STAR, None)], expected_groups=n).run()
        keys = {k[0] for k in got.keys}
        assert len(keys) == ctx.join_table_info(part).rows and all(L.pgf_partition_of_key(k, world) == r for k in keys)
        assert not (keys & seen)
        seen |= keys
        total += len(keys)
        assert abs(len(keys) - n / world) < 0.05 * n / world      # balanced
        ctx.destroy_join_table(part)
    assert total == n
    ctx.destroy_join_table(rows.join_table)
    orders.release()


@pytest.mark.parametrize("world", [1, 4])
def test_partitioned_q3_plan_matches_single_gpu_and_oracle(ctx, world):
    """The plan of tpch.gpu_q3_partitioned with its ranks played in turn: T2_r = rank r's share of the joined orders,
    the lineitem rows the runtime filter lets through are routed to the owner of their key, every group is complete
    on its owner.  The union of the ranks' groups is the single-GPU result (and the oracle's)."""
    ncust, nord, nli = 15_000, 150_000, 600_000
    cust = ctx.gen_scan(GenTable.CUSTOMER_Q3, ncust, seed=42)
    orders = ctx.gen_scan(GenTable.ORDERS_Q3, nord, seed=42, scale_rows=ncust)
    li = ctx.gen_scan(GenTable.LINEITEM_Q3, nli, seed=42, scale_rows=nord)
    single, st1 = T.gpu_q3(ctx, cust, orders, li)
    if world == 1:   # the library path exactly as a 1-rank job runs it
        top, st = T.gpu_q3_partitioned(ctx, cust, orders, li, nord_total=nord, limit=10)
        want10 = U.top10(single)
        assert [(r[0], r[2], r[3]) for r in top] == [(r[0], r[2], r[3]) for r in want10]
        for g, w in zip(top, want10):
            U.assert_close(g[1], w[1], 1e-12, "revenue")
        assert st["final"].rows_out == st1["lineitem"].rows_out and st["lineitem"].rows_out < 0.05 * nli
    else:
        r1 = cust.pipeline().filter(1, Cmp.EQ, b"BUILDING").build_join(0, [], rows_only=True).run()
        t1, _ = ctx.exchange(r1.join_table, partition=False)
        rf = ctx.runtime_filter(T.q3_bloom_params(ncust, nord)[1])
        rf.try_acquire_builder()
        r2 = orders.pipeline().filter(2, Cmp.LT, U.Q3_DATE).join(t1, 1).build_join(0, [2, 3], rf, rows_only=True).run()
        rf.publish_ready()
        r3 = li.pipeline().bloom_probe(rf, 0).filter(3, Cmp.GT, U.Q3_DATE).build_join(0, [1, 2], rows_only=True).run()
        assert r3.rows_out < 0.05 * nli and r3.rows_in - r3.rows_bloom > 0.4 * nli      # the filter does the routing's work
        schema = [ColumnSpec(TypeTag.Int32), ColumnSpec(TypeTag.Float64), ColumnSpec(TypeTag.Float64)]
        union, joined = {}, 0
        for r in range(world):
            t2, _ = ctx.exchange(r2.join_table, partition=True, emulate=(world, r))
            rs3, _ = ctx.exchange(r3.join_table, partition=True, rows_only=True, emulate=(world, r))
            res = (ctx.row_set_pipeline(rs3, schema).join(t2, 0)
                   .aggregate([0, (1, 0), (1, 1)], [(AggFunc.SUM, [Factor.of(1), Factor.const_minus(1.0, 2)])], expected_groups=nord).run())
            joined += res.rows_out
            for k, v in res.by_key().items():
                assert k not in union
                union[k] = v
            ctx.destroy_join_table(t2); ctx.destroy_join_table(rs3)
        assert joined == st1["lineitem"].rows_out
        ref = single.by_key()
        assert set(union) == set(ref)
        for k, v in ref.items():
            U.assert_close(union[k][0], v[0], 1e-12, f"group {k}")
        for h in (r1.join_table, t1, r2.join_table, r3.join_table):
            ctx.destroy_join_table(h)
    tables = [O.OTable.from_pages(s.read_pages(), 65536, U.orc_cols(sc)) for s, sc in
              ((cust, U.CUSTOMER_SCHEMA), (orders, U.ORDERS_SCHEMA), (li, U.LINEITEM_Q3_SCHEMA))]
    want, _ = U.oracle_q3(*tables)
    U.assert_agg_equal(single, want)
    for s in (cust, orders, li):
        s.release()
