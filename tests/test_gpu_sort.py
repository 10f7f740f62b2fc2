"""SortExec / TopK above the aggregate (benches/tpch/queries/q01.sql, q03.sql):
ORDER BY ... [LIMIT n] through the C ABI against a Python sort of the unordered result with
DataFusion's semantics (ASC NULLS LAST, DESC NULLS FIRST, Float64 total order)."""
import numpy as np
import pytest

import pg_fusion_b200 as pg
from pg_fusion_b200 import AggFunc, Cmp, ColumnSpec, Factor, GenTable, TypeTag
from pg_fusion_b200 import arrow_layout as AL

from . import util as U

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = pg.Context()
    yield c
    c.close()


def py_order(rows, terms, limit=0):
    """rows: list of (keys tuple, aggs tuple); terms: (which, index, desc[, nulls_first])."""
    import functools

    def cmp(a, b):
        for t in terms:
            which, index, desc = t[0], t[1], bool(t[2])
            nulls_first = bool(t[3]) if len(t) > 3 else desc
            x = (a[1] if which == "agg" else a[0])[index]
            y = (b[1] if which == "agg" else b[0])[index]
            if x is None or y is None:
                if x is None and y is None:
                    continue
                return -1 if (x is None) == nulls_first else 1
            if x != y:
                c = -1 if x < y else 1
                return -c if desc else c
        return 0
    out = sorted(rows, key=functools.cmp_to_key(cmp))
    return out[:limit] if limit else out


def test_q1_order_by_keys(ctx):
    scan = ctx.gen_scan(GenTable.LINEITEM_Q1, 100_000, seed=42)
    res = U.gpu_q1(scan).order_by([("key", 0, False), ("key", 1, False)]).run()
    assert res.keys == sorted(res.keys) and len(res.keys) == 4
    res = U.gpu_q1(scan).order_by([("agg", 7, True)], limit=2).run()     # ORDER BY count(*) DESC LIMIT 2
    full = U.gpu_q1(scan).run()
    assert [a[7] for a in res.aggs] == sorted((a[7] for a in full.aggs), reverse=True)[:2]
    scan.release()


@pytest.mark.parametrize("limit", [10, 64, 80])
def test_q3_top_k_by_revenue(ctx, limit):
    """ORDER BY revenue DESC, o_orderdate LIMIT n: n <= 64 selects on the device, above on the host."""
    ncust, nord, nli = 1500, 15_000, 60_000
    customer = ctx.gen_scan(GenTable.CUSTOMER_Q3, ncust, seed=42)
    orders = ctx.gen_scan(GenTable.ORDERS_Q3, nord, seed=42, scale_rows=ncust)
    lineitem = ctx.gen_scan(GenTable.LINEITEM_Q3, nli, seed=42, scale_rows=nord)
    r1 = customer.pipeline().filter(1, Cmp.EQ, b"BUILDING").build_join(0, []).run()
    r2 = orders.pipeline().filter(2, Cmp.LT, U.Q3_DATE).join(r1.join_table, 1).build_join(0, [2, 3]).run()

    def plan():
        return (lineitem.pipeline().filter(3, Cmp.GT, U.Q3_DATE).join(r2.join_table, 0)
                .aggregate([0, (1, 0), (1, 1)], [(AggFunc.SUM, [Factor.of(1), Factor.const_minus(1.0, 2)])],
                           expected_groups=max(1024, r2.rows_out)))
    full = plan().run()
    assert len(full.keys) > limit
    terms = [("agg", 0, True), ("key", 1, False), ("key", 0, False)]
    got = plan().order_by(terms, limit=limit).run()
    want = py_order(list(zip(full.keys, full.aggs)), terms, limit)
    assert len(got.keys) == limit
    # Float64 sums are not bit-reproducible run to run: compare the order by key and the sums with tolerance
    assert got.keys == [w[0] for w in want]
    for a, w in zip(got.aggs, want):
        U.assert_close(a[0], w[1][0], 1e-12, "revenue")
    ctx.destroy_join_table(r1.join_table)
    ctx.destroy_join_table(r2.join_table)
    for s in (customer, orders, lineitem):
        s.release()


@pytest.mark.parametrize("limit,nkeys", [(7, 300), (0, 300), (7, 4000), (64, 4000)])
def test_order_by_with_null_keys_ties_and_integer_sums(ctx, limit, nkeys):
    """(4000 keys x 7 strings = ~28 000 groups: the device selection runs two levels; 300 keys: one.)"""
    r = np.random.default_rng(11)
    n = 120_000
    schema = [ColumnSpec(TypeTag.Int32, True), ColumnSpec(TypeTag.Utf8View, True), ColumnSpec(TypeTag.Int64, True)]
    k = r.integers(0, nkeys, n).astype(np.int32)
    s = [bytes([66 + int(x)]) * int(1 + x % 3) for x in r.integers(0, 6, n)]
    v = r.integers(-5, 6, n)          # small values: many tied sums
    valid = [r.random(n) > 0.05, r.random(n) > 0.1, r.random(n) > 0.2]
    pages = AL.encode_pages(schema, [(k, valid[0]), (AL.inline_views(s), valid[1]), (v, valid[2])], rows_per_page=900)
    scan = ctx.declare_scan(schema)
    scan.push_pages(pages)
    scan.finish()

    def plan():
        return scan.pipeline().aggregate([0, 1], [(AggFunc.SUM, [Factor.of(2)]), (AggFunc.COUNT, [Factor.of(2)]), (AggFunc.COUNT_STAR, None)])
    full = plan().run()
    rows = list(zip(full.keys, full.aggs))
    for terms in ([("agg", 0, True), ("key", 1, False), ("key", 0, True)],                  # sum DESC (NULLs first), s ASC (NULLs last), k DESC
                  [("key", 1, False, True), ("agg", 2, False), ("key", 0, False)],          # s ASC NULLS FIRST, count(*) ASC, k ASC
                  [("agg", 1, True), ("key", 0, False, True), ("key", 1, True, False)],
                  [("key", 0, True, False), ("key", 1, False, True)]):                        # k DESC NULLS LAST: the first term alone leaves ties
        got = plan().order_by(terms, limit=limit).run()
        want = py_order(rows, terms, limit)
        assert list(zip(got.keys, got.aggs)) == want   # integer sums and counts: exact, total order given all keys
    scan.release()


def test_order_by_terms_are_validated_by_the_eligibility_check(ctx):
    """A malformed ORDER BY (index outside the output columns, too many terms) is an argument error of
    pgf_pipeline_check and of every run / merge entry point -- never an out-of-bounds read."""
    scan = ctx.gen_scan(GenTable.LINEITEM_Q1, 10_000, seed=1)
    for terms in ([("key", 2, False)], [("agg", 8, True)], [("key", -1, False)]):
        p = U.gpu_q1(scan).order_by(terms)
        assert p.check() == 1                      # PGF_ERR_INVALID_ARGUMENT
        with pytest.raises(pg.PgfError) as e:
            p.run()
        assert e.value.code == 1
    p = U.gpu_q1(scan).order_by([("key", 0, False)])
    p.p.nsort = 5                                  # > PGF_MAX_SORT
    assert p.check() == 1
    assert U.gpu_q1(scan).order_by([("key", 1, True), ("agg", 7, False)]).check() == 0
    scan.release()
