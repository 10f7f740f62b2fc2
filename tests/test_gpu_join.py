"""GPU parity tests of the int-key hash join (build + probe), the Bloom filter fused into
pipelines and the TPC-H Q3 shape, against the oracle.  Join row multisets are bit-exact."""
import numpy as np
import pytest

import pg_fusion_b200 as pg
from oracle import pyorc as O
from pg_fusion_b200 import AggFunc, BloomParams, Cmp, ColumnSpec, Factor, GenTable, TypeTag
from pg_fusion_b200 import arrow_layout as AL

from . import util as U

pytestmark = pytest.mark.gpu
E = O.Expr


@pytest.fixture(scope="module")
def ctx():
    c = pg.Context(keep_redundant_bloom_probes=True)   # these tests check the fused Bloom probe itself
    yield c
    c.close()


def load(ctx, schema, cols, **kw):
    pages = AL.encode_pages(schema, cols, **kw)
    scan = ctx.declare_scan(schema)
    scan.push_pages(pages)
    scan.finish()
    return scan, O.OTable.from_pages(pages, 65536, U.orc_cols(schema))


def test_reference_smoke_join(ctx):
    """pg/extension/src/smoke_tests.rs:253-304: {1,2,3} |><| {2,3,4} on bigint keys = 2 rows, score 20 for key 2."""
    s1 = [ColumnSpec(TypeTag.Int64), ColumnSpec(TypeTag.Int64)]
    left, _ = load(ctx, s1, [(np.array([1, 2, 3], np.int64), None), (np.array([10, 20, 30], np.int64), None)])
    right, _ = load(ctx, [ColumnSpec(TypeTag.Int64)], [(np.array([2, 3, 4], np.int64), None)])
    b = left.pipeline().build_join(0, [1]).run()
    assert b.rows_out == 3
    r = right.pipeline().join(b.join_table, 0).aggregate([0], [(AggFunc.COUNT_STAR, None), (AggFunc.SUM, [Factor.of((1, 0))])]).run()
    assert r.rows_out == 2
    assert r.by_key() == {(2,): (1, 20), (3,): (1, 30)}


@pytest.mark.parametrize("key_dtype,tag", [(np.int16, TypeTag.Int16), (np.int32, TypeTag.Int32), (np.int64, TypeTag.Int64)])
def test_join_multiset_with_duplicates_and_nulls(ctx, key_dtype, tag):
    r = np.random.default_rng(4)
    nb, npr = 5000, 40_000
    bk = r.integers(0, 2000, nb).astype(key_dtype)           # duplicates on the build side
    pk = r.integers(-500, 2500, npr).astype(key_dtype)       # misses and duplicates on the probe side
    bvalid, pvalid = r.random(nb) > 0.05, r.random(npr) > 0.05
    bpay = r.integers(-10**6, 10**6, nb).astype(np.int64)
    pval = r.integers(-10**6, 10**6, npr).astype(np.int64)
    bs = [ColumnSpec(tag, True), ColumnSpec(TypeTag.Int64)]
    ps = [ColumnSpec(tag, True), ColumnSpec(TypeTag.Int64)]
    build, bt = load(ctx, bs, [(bk, bvalid), (bpay, None)], rows_per_page=900)
    probe, pt = load(ctx, ps, [(pk, pvalid), (pval, None)], rows_per_page=1100)
    b = build.pipeline().build_join(0, [1]).run()
    assert b.rows_out == nb
    # number of matched pairs == oracle HashJoinExec pair count (NULL keys never match)
    want_b, want_p = O.hash_join_pairs(bt, 0, pt, 0)
    got = probe.pipeline().join(b.join_table, 0).count().run()
    assert got.rows_out == want_b.size
    # multiset check via aggregates over (build payload, probe value) grouped by key:
    # sum(build.pay), sum(probe.val), count(*) are bit-exact iff the pair multiset per key matches
    res = (probe.pipeline().join(b.join_table, 0)
           .aggregate([0], [(AggFunc.SUM, [Factor.of((1, 0))]), (AggFunc.SUM, [Factor.of(1)]), (AggFunc.COUNT_STAR, None)]).run())
    want = O.aggregate(pt, None, [E.col(0)], [(O.AGG_SUM, E.col(1, 1)), (O.AGG_SUM, E.col(1)), (O.AGG_COUNT_STAR, None)],
                       joins=[(bt, 0, 0, 0)])
    U.assert_agg_equal(res, want, rel=0)
    # a filter on the probe side and one on the build side
    b2 = build.pipeline().filter(1, Cmp.GE, 0).build_join(0, [1]).run()
    res = (probe.pipeline().filter(1, Cmp.LT, 0).join(b2.join_table, 0)
           .aggregate([], [(AggFunc.SUM, [Factor.of((1, 0))]), (AggFunc.COUNT_STAR, None)]).run())
    bt_f = bt.select(O.filter_rows(bt, E.col(1).ge(E.i64(0))))
    want = O.aggregate(pt, E.col(1).lt(E.i64(0)), [], [(O.AGG_SUM, E.col(1, 1)), (O.AGG_COUNT_STAR, None)], joins=[(bt_f, 0, 0, 0)])
    U.assert_agg_equal(res, want, rel=0)
    ctx.destroy_join_table(b.join_table)
    ctx.destroy_join_table(b2.join_table)


def test_empty_build_and_empty_probe(ctx):
    s = [ColumnSpec(TypeTag.Int32)]
    empty = ctx.declare_scan(s); empty.finish()
    some, _ = load(ctx, s, [(np.arange(100, dtype=np.int32), None)])
    b = empty.pipeline().build_join(0).run()
    assert b.rows_out == 0
    assert some.pipeline().join(b.join_table, 0).count().run().rows_out == 0
    b2 = some.pipeline().build_join(0).run()
    assert empty.pipeline().join(b2.join_table, 0).count().run().rows_out == 0
    assert some.pipeline().join(b2.join_table, 0).count().run().rows_out == 100


def test_bloom_fused_into_build_and_probe_pipelines(ctx):
    """RuntimeFilterBuildExec semantics: the build pipeline populates the filter with every non-null
    build key (bit-exact vs the oracle); probing it in the scan pipeline never changes join results."""
    r = np.random.default_rng(12)
    bk = r.integers(0, 10**6, 20_000).astype(np.int32)
    pk = r.integers(0, 10**6, 100_000).astype(np.int32)
    bvalid = r.random(bk.size) > 0.1
    build, bt = load(ctx, [ColumnSpec(TypeTag.Int32, True)], [(bk, bvalid)])
    probe, pt = load(ctx, [ColumnSpec(TypeTag.Int32)], [(pk, None)])
    p = BloomParams.new(**pg.GUC_DEFAULT_BLOOM)
    rf = ctx.runtime_filter(p)
    rf.try_acquire_builder()
    b = build.pipeline().build_join(0, [], rf).run()
    assert b.bloom_rows == int(bvalid.sum())      # RuntimeFilterBuildRowsTotal
    rf.publish_ready()
    ob = O.Bloom(O.bloom_params(p.bit_count, p.hash_count, p.seed))
    ob.insert_keys(bk, np.packbits(bvalid, bitorder="little"))
    assert (rf.words() == ob.words).all()
    plain = probe.pipeline().join(b.join_table, 0).count().run()
    filt = probe.pipeline().bloom_probe(rf, 0).join(b.join_table, 0).count().run()
    keep, rejected = ob.probe_keys(pk)
    assert filt.rows_bloom == int(keep.sum()) and filt.rows_in - filt.rows_bloom == rejected
    assert filt.rows_out == plain.rows_out == O.hash_join_pairs(bt, 0, pt, 0)[0].size
    # a filter that is not Ready (stale generation) must pass rows unfiltered
    stale = probe.pipeline().bloom_probe(rf, 0, generation=rf.generation + 1).count().run()
    assert stale.rows_bloom == stale.rows_in


@pytest.mark.parametrize("with_bloom", [False, True])
def test_q3_shape_matches_oracle(ctx, with_bloom):
    ncust, nord, nli = 1500, 15_000, 60_000
    customer = ctx.gen_scan(GenTable.CUSTOMER_Q3, ncust, seed=42)
    orders = ctx.gen_scan(GenTable.ORDERS_Q3, nord, seed=42, scale_rows=ncust)
    lineitem = ctx.gen_scan(GenTable.LINEITEM_Q3, nli, seed=42, scale_rows=nord)
    ct = O.OTable.from_pages(customer.read_pages(), 65536, U.orc_cols(U.CUSTOMER_SCHEMA))
    ot = O.OTable.from_pages(orders.read_pages(), 65536, U.orc_cols(U.ORDERS_SCHEMA))
    lt = O.OTable.from_pages(lineitem.read_pages(), 65536, U.orc_cols(U.LINEITEM_Q3_SCHEMA))
    want, wstats = U.oracle_q3(ct, ot, lt)
    bp = (BloomParams.new(**pg.GUC_DEFAULT_BLOOM), BloomParams.new(1 << 16, 4, 7)) if with_bloom else None
    res, stats = U.gpu_q3(ctx, customer, orders, lineitem, bp)
    assert stats["customer"].rows_out == wstats["customers"]
    assert stats["orders"].rows_out == wstats["orders"]
    assert res.rows_out == want.rows_joined and len(res.keys) == len(want.keys) > 0
    U.assert_agg_equal(res, want)
    got10, want10 = U.top10(res), U.top10(want)
    assert [(r[0], r[2], r[3]) for r in got10] == [(r[0], r[2], r[3]) for r in want10]
    for g, w in zip(got10, want10):
        U.assert_close(g[1], w[1], 1e-12, "revenue")
    if with_bloom:
        assert stats["lineitem"].rows_bloom < stats["lineitem"].rows_in   # the filter does reject rows
    for s in (customer, orders, lineitem):
        s.release()


def test_redundant_and_saturated_bloom_probes_are_dropped_by_default():
    """A runtime filter is an optimisation only.  By default a fused probe is dropped when the same
    pipeline probes the join table on that key, or when the filter is saturated (fill^k > 0.9): the
    result is identical, rows_bloom == rows_in shows that no row was probed."""
    with pg.Context() as c:
        r = np.random.default_rng(5)
        bk = r.integers(0, 10**6, 20_000).astype(np.int32)
        pk = r.integers(0, 10**6, 100_000).astype(np.int32)
        build, bt = load(c, [ColumnSpec(TypeTag.Int32)], [(bk, None)])
        probe, pt = load(c, [ColumnSpec(TypeTag.Int32)], [(pk, None)])
        rf = c.runtime_filter(BloomParams.new(1 << 20, 4, 7))            # sparse: fill ~ 7 %
        rf.try_acquire_builder()
        b = build.pipeline().build_join(0, [], rf).run()
        rf.publish_ready()
        want = O.hash_join_pairs(bt, 0, pt, 0)[0].size
        joined = probe.pipeline().bloom_probe(rf, 0).join(b.join_table, 0).count().run()
        assert joined.rows_out == want and joined.rows_bloom == joined.rows_in      # (a) redundant next to the join probe
        alone = probe.pipeline().bloom_probe(rf, 0).count().run()
        assert alone.rows_bloom < alone.rows_in and alone.rows_out == alone.rows_bloom   # kept when it is the only filter
        sat = c.runtime_filter(BloomParams.new(1 << 12, 4, 7))           # 20 000 keys in 4096 bits: saturated
        sat.try_acquire_builder()
        sat.insert_keys(bk)
        sat.publish_ready()
        dropped = probe.pipeline().bloom_probe(sat, 0).count().run()
        assert dropped.rows_bloom == dropped.rows_in == dropped.rows_out                # (b) saturated


def test_two_join_probes_on_one_stream(ctx):
    """VERDICT r1 grammar gap: two HashJoinExec probes fused into one scan stream -- lineitem probes the orders
    table and, with the matched order's o_custkey (a payload of the first join), the customer table.  Checked
    against the oracle's nested hash joins (multisets: duplicates on both build sides multiply)."""
    pages, (ct, ot, lt) = U.q3_host_tables(300, 3000, 40_000, seed=21, dup_keys=True, rows_per_page=700)
    scans = []
    for pg_, schema in zip(pages, (U.CUSTOMER_SCHEMA, U.ORDERS_SCHEMA, U.LINEITEM_Q3_SCHEMA)):
        s = ctx.declare_scan(schema)
        s.push_pages(pg_)
        s.finish()
        scans.append(s)
    cust, orders, li = scans
    tc = cust.pipeline().filter(1, Cmp.EQ, b"BUILDING").build_join(0, []).run()
    to = orders.pipeline().filter(2, Cmp.LT, U.Q3_DATE).build_join(0, [1, 3]).run()      # payload: o_custkey, o_shippriority
    probe = lambda: li.pipeline().filter(3, Cmp.GT, U.Q3_DATE).join(to.join_table, 0).join(tc.join_table, (1, 0))
    res = probe().aggregate([0], [(AggFunc.COUNT_STAR, None), (AggFunc.SUM, [Factor.of(1)])], expected_groups=4096).run()
    res_i = probe().aggregate([0], [(AggFunc.SUM, [Factor.of((1, 1))])], expected_groups=4096).run()   # Int64 sums: exact
    assert res.variant == "compact_1_string_term"
    # oracle: orders |><| customer first (as a table), then lineitem |><| that
    cust_f = ct.select(O.filter_rows(ct, E.col(1).eq(E.s(b"BUILDING"))))
    ord_f = ot.select(O.filter_rows(ot, E.col(2).lt(E.s(U.Q3_DATE))))
    _, probe_rows = O.hash_join_pairs(cust_f, 0, ord_f, 1)
    ord_j = ord_f.take(probe_rows)
    want = O.aggregate(lt, E.col(3).gt(E.s(U.Q3_DATE)), [E.col(0)], [(O.AGG_COUNT_STAR, None), (O.AGG_SUM, E.col(1))], joins=[(ord_j, 0, 0, 0)])
    want_i = O.aggregate(lt, E.col(3).gt(E.s(U.Q3_DATE)), [E.col(0)], [(O.AGG_SUM, E.col(3, 1))], joins=[(ord_j, 0, 0, 0)])
    assert res.rows_out == res_i.rows_out == want.rows_joined > 0
    U.assert_agg_equal(res, want, rel=1e-12)
    U.assert_agg_equal(res_i, want_i, rel=0)
    for h in (tc.join_table, to.join_table):
        ctx.destroy_join_table(h)
    for s in scans:
        s.release()


def test_split_execution_and_its_fused_fallback_agree(ctx, monkeypatch):
    """Aggregates behind one join run as two kernels (stages A + B, then stage C over the tag hits).  The same plan as
    one fused kernel (PGF_PROBE_SPLIT=0) and with an entry buffer that is far too small (the run is repeated fused)
    must return the same groups, counts and integer sums; duplicates on the build side multiply in all three."""
    r = np.random.default_rng(12)
    nb, npr = 20_000, 300_000
    bk = r.integers(0, 8000, nb).astype(np.int32)
    pk = r.integers(-1000, 12_000, npr).astype(np.int32)
    bpay = r.integers(-10**6, 10**6, nb).astype(np.int64)
    pval = r.integers(-10**6, 10**6, npr).astype(np.int64)
    bs = [ColumnSpec(TypeTag.Int32), ColumnSpec(TypeTag.Int64)]
    build, bt = load(ctx, bs, [(bk, None), (bpay, None)])
    probe, pt = load(ctx, bs, [(pk, None), (pval, None)])
    b = build.pipeline().build_join(0, [1]).run()

    def run():
        return (probe.pipeline().filter(1, Cmp.GE, -500_000).join(b.join_table, 0)
                .aggregate([0], [(AggFunc.SUM, [Factor.of((1, 0))]), (AggFunc.SUM, [Factor.of(1)]), (AggFunc.COUNT_STAR, None)],
                           expected_groups=16_000).run())   # (room for every group: no table-overflow re-run blurs the launch counts)
    split = run()
    monkeypatch.setenv("PGF_PROBE_SPLIT", "0")
    fused = run()
    monkeypatch.delenv("PGF_PROBE_SPLIT")
    monkeypatch.setenv("PGF_PROBE_SPLIT_CAP", "64")
    fallback = run()
    monkeypatch.delenv("PGF_PROBE_SPLIT_CAP")
    want = O.aggregate(pt, E.col(1).ge(E.i64(-500_000)), [E.col(0)], [(O.AGG_SUM, E.col(1, 1)), (O.AGG_SUM, E.col(1)), (O.AGG_COUNT_STAR, None)],
                       joins=[(bt, 0, 0, 0)])
    for res in (split, fused, fallback):
        assert res.rows_out == split.rows_out
        U.assert_agg_equal(res, want, rel=0)
    assert split.kernel_launches == fused.kernel_launches + 1          # scan kernel + stage-C kernel
    assert fallback.kernel_launches > split.kernel_launches            # the split attempt, then the fused run
    build.release()
    probe.release()
