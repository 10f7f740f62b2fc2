"""Cross-checks the operator oracle (oracle/orc_ops.c, [DF-K] restatement of DataFusion 44
semantics) against pyarrow / Acero -- an independent columnar engine (Arrow C++), NOT
DataFusion -- and against the reference's PostgreSQL-bound smoke values.  CPU only."""
import decimal

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

from oracle import pyorc as O
from pg_fusion_b200 import ColumnSpec, TypeTag
from pg_fusion_b200 import arrow_layout as AL

from . import util as U

E = O.Expr


def table_of(schema, cols, **kw):
    return O.OTable.from_pages(AL.encode_pages(schema, cols, **kw), 65536, U.orc_cols(schema))


def test_q6_against_acero():
    li = U.lineitem(30_000, 1)
    t = O.OTable.from_pages(U.q6_pages(li), 65536, U.orc_cols(U.Q6_SCHEMA))
    got = U.oracle_q6(t)
    dates = pa.array([d.decode() for d in U.dates_from_days(li["ship"])])
    tb = pa.table({"q": li["qty"], "p": li["price"], "d": li["disc"], "s": dates})
    m = pc.and_(pc.and_(pc.greater_equal(tb["s"], "1994-01-01"), pc.less(tb["s"], "1995-01-01")),
                pc.and_(pc.and_(pc.greater_equal(tb["d"], 0.05), pc.less_equal(tb["d"], 0.07)), pc.less(tb["q"], 24.0)))
    f = tb.filter(m)
    assert got.rows_filtered == f.num_rows
    want = pc.sum(pc.multiply(f["p"], f["d"])).as_py()
    U.assert_close(got.aggs[0][0], want, 1e-12, "q6 revenue")
    assert got.aggs[0][1] == f.num_rows


def test_q1_against_acero():
    li = U.lineitem(30_000, 2)
    t = O.OTable.from_pages(U.q1_pages(li), 65536, U.orc_cols(U.Q1_SCHEMA))
    got = U.oracle_q1(t).by_key()
    tb = pa.table({"q": li["qty"], "p": li["price"], "d": li["disc"], "t": li["tax"],
                   "rf": [bytes(x).decode() for x in li["rf"]], "ls": [bytes(x).decode() for x in li["ls"]],
                   "s": [d.decode() for d in U.dates_from_days(li["ship"])]})
    tb = tb.filter(pc.less_equal(tb["s"], "1998-09-02"))
    dp = pc.multiply(tb["p"], pc.subtract(1.0, tb["d"]))
    tb = tb.append_column("dp", dp).append_column("ch", pc.multiply(dp, pc.add(1.0, tb["t"])))
    g = tb.group_by(["rf", "ls"]).aggregate([("q", "sum"), ("p", "sum"), ("dp", "sum"), ("ch", "sum"), ("q", "mean"),
                                             ("p", "mean"), ("d", "mean"), ("q", "count")]).to_pylist()
    assert len(g) == len(got) == 4
    for row in g:
        a = got[(row["rf"].encode(), row["ls"].encode())]
        for x, y in zip(a, (row["q_sum"], row["p_sum"], row["dp_sum"], row["ch_sum"], row["q_mean"], row["p_mean"], row["d_mean"])):
            U.assert_close(x, y, 1e-11, "q1")   # Acero sums pairwise, the oracle sequentially
        assert a[7] == row["q_count"]


def test_nulls_int_sums_and_join_against_acero():
    r = np.random.default_rng(5)
    n = 20_000
    k = r.integers(0, 50, n).astype(np.int32)
    v = r.integers(-2**40, 2**40, n)
    kv, vv = r.random(n) > 0.1, r.random(n) > 0.2
    schema = [ColumnSpec(TypeTag.Int32, True), ColumnSpec(TypeTag.Int64, True)]
    t = table_of(schema, [(k, kv), (v, vv)], rows_per_page=999)
    got = O.aggregate(t, E.col(1).gt(E.i64(0)), [E.col(0)], [(O.AGG_SUM, E.col(1)), (O.AGG_COUNT, E.col(1)), (O.AGG_COUNT_STAR, None)]).by_key()
    tb = pa.table({"k": pa.array(k, mask=~kv), "v": pa.array(v, mask=~vv)})
    f = tb.filter(pc.greater(tb["v"], 0))     # NULL predicate rows are dropped
    g = f.group_by("k").aggregate([("v", "sum"), ("v", "count"), ([], "count_all")]).to_pylist()
    assert len(g) == len(got)
    for row in g:
        assert got[(row["k"],)] == (row["v_sum"], row["v_count"], row["count_all"])
    # inner join multiset size; NULL keys never match; duplicates multiply
    bk = r.integers(0, 300, 2000).astype(np.int32)
    bt = table_of([ColumnSpec(TypeTag.Int32, True)], [(bk, r.random(2000) > 0.1)])
    b, p = O.hash_join_pairs(bt, 0, t, 0)
    bcol, bvalid = bt.column(0)
    left = pa.table({"k": pa.array(bcol, mask=None if bvalid is None else ~bvalid.astype(bool))})
    j = left.join(tb, keys="k", join_type="inner")
    assert b.size == j.num_rows
    # probe order is preserved and build matches come in build order
    assert (np.diff(p.astype(np.int64)) >= 0).all()


def test_decimal128_sums_against_python_and_acero():
    r = np.random.default_rng(6)
    n = 5000
    price = r.integers(0, 10**12, n)
    disc = r.integers(0, 11, n)

    def dec(vals):
        out = np.zeros((len(vals), 16), dtype=np.uint8)
        for i, v in enumerate(vals):
            out[i] = np.frombuffer((int(v) & (2**128 - 1)).to_bytes(16, "little"), dtype=np.uint8)
        return out
    schema = [ColumnSpec(TypeTag.Decimal128), ColumnSpec(TypeTag.Decimal128)]
    t = table_of(schema, [(dec(price), None), (dec(disc), None)])
    got = O.aggregate(t, None, [], [(O.AGG_SUM, E.col(0) * (E.i128(100) - E.col(1))), (O.AGG_SUM, E.col(0))])
    want = sum(int(p) * (100 - int(d)) for p, d in zip(price, disc))
    assert got.aggs[0] == (want, int(price.sum()))
    # the unscaled i128 is what Arrow's decimal128 arithmetic produces too (declared precision differs)
    a = pa.array([decimal.Decimal(int(p)).scaleb(-2) for p in price], type=pa.decimal128(15, 2))
    d = pa.array([decimal.Decimal(int(x)).scaleb(-2) for x in disc], type=pa.decimal128(15, 2))
    one = pa.scalar(decimal.Decimal("1.00"), type=pa.decimal128(15, 2))
    s = pc.sum(pc.multiply(a, pc.subtract(one, d))).as_py()
    assert int(s.scaleb(4)) == want


def test_reference_smoke_values_on_the_oracle():
    # pg/extension/src/smoke_tests.rs:205-251
    ids = np.arange(1, 50_001, dtype=np.int64)
    t = table_of([ColumnSpec(TypeTag.Int64)], [(ids, None)])
    assert O.aggregate(t, None, [], [(O.AGG_AVG, E.col(0))], sum_lanes=4).aggs[0][0] == 25000.5
    ids = np.arange(1, 5001, dtype=np.int64)
    t = table_of([ColumnSpec(TypeTag.Int64)], [(ids, None)])
    assert O.aggregate(t, None, [], [(O.AGG_COUNT, E.col(0)), (O.AGG_SUM, E.col(0))]).aggs[0] == (5000, 12502500)
    # smoke_tests.rs:253-304
    a = table_of([ColumnSpec(TypeTag.Int64), ColumnSpec(TypeTag.Int64)], [(np.array([1, 2, 3]), None), (np.array([10, 20, 30]), None)])
    b = table_of([ColumnSpec(TypeTag.Int64)], [(np.array([2, 3, 4]), None)])
    bi, pi = O.hash_join_pairs(a, 0, b, 0)
    assert bi.size == 2 and sorted(bi.tolist()) == [1, 2]


def test_float_total_order_and_kleene_and():
    # arrow compares floats by totalOrder: -0.0 < +0.0, NaN is greatest and equal to itself [DF-K]
    vals = np.array([-0.0, 0.0, np.nan, 1.0, -np.inf])
    t = table_of([ColumnSpec(TypeTag.Float64)], [(vals, None)])
    assert O.filter_rows(t, E.col(0).lt(E.f64(0.0))).tolist() == [1, 0, 0, 0, 1]
    assert O.filter_rows(t, E.col(0).eq(E.f64(float("nan")))).tolist() == [0, 0, 1, 0, 0]
    assert O.filter_rows(t, E.col(0).gt(E.f64(1e308))).tolist() == [0, 0, 1, 0, 0]
    # FALSE AND NULL = FALSE, TRUE AND NULL = NULL (dropped either way, but COUNT over a
    # filtered aggregate must not see the row)
    schema = [ColumnSpec(TypeTag.Int32, True), ColumnSpec(TypeTag.Int32, True)]
    t = table_of(schema, [(np.array([1, 1, 5], np.int32), np.array([True, False, True])),
                          (np.array([1, 1, 1], np.int32), np.array([True, True, False]))])
    f = E.col(0).eq(E.i64(1)).and_(E.col(1).eq(E.i64(1)))
    assert O.filter_rows(t, f).tolist() == [1, 0, 0]


@pytest.mark.parametrize("dup_keys,nthreads", [(False, 1), (True, 1), (True, 4)])
def test_q3_stream_oracle_equals_the_generic_interpreter(dup_keys, nthreads):
    """oracle/orc_q3.c (the page-sharded Q3 loops used for full-size parity and as the CPU baseline) against the
    generic operator interpreter (orc_ops.c: HashJoinExec pairs + AggregateExec), fed in several shards."""
    pages, tables = U.q3_host_tables(400, 3000, 20_000, seed=11, dup_keys=dup_keys, rows_per_page=500)
    want, info = U.oracle_q3(*tables)
    q = O.Q3Stream(nthreads=nthreads)
    for name, pg_ in zip(("customer", "orders", "lineitem"), pages):
        half = pg_.shape[0] // 2
        getattr(q, name)(pg_[:half])
        getattr(q, name)(pg_[half:])
    st = q.stats()
    assert st["customers"] == info["customers"] and st["orders"] == info["orders"]
    assert st["lineitem_rows"] == 20_000 and st["filtered"] == want.rows_filtered and st["joined"] == want.rows_joined
    U.assert_q3_stream_equals(want, q.groups(), rel=0 if nthreads == 1 and not dup_keys else 1e-13)
    q.close()
