"""GPU parity tests of the fused scan pipelines (filter / projection / aggregate) through
the C ABI against the oracle (oracle/orc_ops.c) on the same seeded pages.
Bar: counts, integer and Decimal128 sums bit-exact; Float64 SUM/AVG within 1e-12 relative
(BASELINE.json north_star)."""
import numpy as np
import pytest

import pg_fusion_b200 as pg
from oracle import pyorc as O
from pg_fusion_b200 import AggFunc, Cmp, ColumnSpec, Factor, GenTable, TypeTag
from pg_fusion_b200 import arrow_layout as AL

from . import util as U

pytestmark = pytest.mark.gpu
E = O.Expr


@pytest.fixture(scope="module")
def ctx():
    c = pg.Context()
    yield c
    c.close()


def load(ctx, schema, pages):
    scan = ctx.declare_scan(schema)
    if len(pages):
        scan.push_pages(pages)
    scan.finish()
    return scan


@pytest.mark.parametrize("n,rows_per_page", [(0, None), (1, None), (999, None), (50_000, None), (20_000, 100), (3000, 1)])
def test_q6_shape_matches_oracle(ctx, n, rows_per_page):
    li = U.lineitem(n, 11 + n)
    pages = U.q6_pages(li, rows_per_page=rows_per_page) if n else np.zeros((0, 65536), np.uint8)
    scan = load(ctx, U.Q6_SCHEMA, pages)
    res = U.gpu_q6(scan).run()
    table = O.OTable.from_pages(pages, 65536, U.orc_cols(U.Q6_SCHEMA)) if n else None
    if n == 0:
        assert res.rows_in == 0 and res.aggs == [(None, 0)]  # SUM over no rows is NULL, COUNT(*) = 0
        return
    want = U.oracle_q6(table)
    assert res.rows_in == n and res.rows_filtered == want.rows_filtered
    U.assert_agg_equal(res, want)
    # the lane-striped "arrow sum kernel" order must agree within the same tolerance
    U.assert_agg_equal(res, U.oracle_q6(table, sum_lanes=4))
    # fast CPU baseline loop agrees with the generic interpreter
    s, rows_in, kept = O.q6_pages(pages, 65536, 2)
    assert (rows_in, kept) == (n, want.rows_filtered)
    if kept:
        U.assert_close(s, want.aggs[0][0], 1e-12, "orc_q6_pages")
    else:
        assert want.aggs[0][0] is None and res.aggs[0][0] is None  # SUM over no rows is NULL
    scan.release()


@pytest.mark.parametrize("n", [5, 4000, 60_000])
def test_q1_shape_matches_oracle(ctx, n):
    li = U.lineitem(n, 3 + n)
    pages = U.q1_pages(li)
    scan = load(ctx, U.Q1_SCHEMA, pages)
    res = U.gpu_q1(scan).run()
    want = U.oracle_q1(O.OTable.from_pages(pages, 65536, U.orc_cols(U.Q1_SCHEMA)))
    assert res.rows_in == n and res.rows_filtered == want.rows_filtered
    U.assert_agg_equal(res, want)
    fast, _ = O.q1_pages(pages, 65536, 3)
    for k, a in want.by_key().items():
        g = fast[(k[0], k[1])]
        U.assert_close(g["sum_charge"], a[3], 1e-12, "orc_q1_pages")
        assert g["count"] == a[7]
    scan.release()


def test_generated_pages_are_valid_and_match_oracle(ctx):
    """The device generator writes reference-format pages; the oracle decodes the same bytes."""
    for table, schema, run_gpu, run_orc in (
            (GenTable.LINEITEM_Q6, U.Q6_SCHEMA, U.gpu_q6, U.oracle_q6),
            (GenTable.LINEITEM_Q1, U.Q1_SCHEMA, U.gpu_q1, U.oracle_q1)):
        scan = ctx.gen_scan(table, 30_000, seed=42)
        pages = scan.read_pages()
        for p in pages[:2]:
            assert AL.import_check(0x4152, 0, p[20:], schema) == 0
            assert O.import_check(0x4152, 0, np.ascontiguousarray(p[20:]), U.orc_cols(schema)) == 0
        want = run_orc(O.OTable.from_pages(pages, 65536, U.orc_cols(schema)))
        res = run_gpu(scan).run()
        assert res.rows_in == 30_000
        U.assert_agg_equal(res, want)
        if table == GenTable.LINEITEM_Q1:
            assert sorted(res.by_key()) == [(b"A", b"F"), (b"N", b"F"), (b"N", b"O"), (b"R", b"F")]
        scan.release()


def test_nulls_three_valued_logic_and_null_groups(ctx):
    r = np.random.default_rng(1)
    n = 20_000
    schema = [ColumnSpec(TypeTag.Int32, True), ColumnSpec(TypeTag.Float64, True), ColumnSpec(TypeTag.Float64, True),
              ColumnSpec(TypeTag.Utf8View, True), ColumnSpec(TypeTag.Int64, True)]
    g = r.integers(0, 3, n).astype(np.int32)
    a = r.random(n) * 100
    b = r.random(n)
    s = [bytes([65 + int(x)]) * int(1 + x) for x in r.integers(0, 5, n)]
    i = r.integers(-10**12, 10**12, n)
    valid = [r.random(n) > p for p in (0.1, 0.2, 0.15, 0.1, 0.3)]
    cols = [(g, valid[0]), (a, valid[1]), (b, valid[2]), (AL.inline_views(s), valid[3]), (i, valid[4])]
    pages = AL.encode_pages(schema, cols, rows_per_page=700)
    scan = load(ctx, schema, pages)
    table = O.OTable.from_pages(pages, 65536, U.orc_cols(schema))
    # WHERE a > 10 AND s >= 'BB' GROUP BY g : sum(a*b), avg(a), count(b), count(*), sum(a*(1-b))
    res = (scan.pipeline().filter(1, Cmp.GT, 10.0).filter(3, Cmp.GE, b"BB")
           .aggregate([0], [(AggFunc.SUM, [Factor.of(1), Factor.of(2)]), (AggFunc.AVG, [Factor.of(1)]),
                            (AggFunc.COUNT, [Factor.of(2)]), (AggFunc.COUNT_STAR, None),
                            (AggFunc.SUM, [Factor.of(1), Factor.const_minus(1.0, 2)])]).run())
    filt = E.col(1).gt(E.f64(10.0)).and_(E.col(3).ge(E.s(b"BB")))
    want = O.aggregate(table, filt, [E.col(0)],
                       [(O.AGG_SUM, E.col(1) * E.col(2)), (O.AGG_AVG, E.col(1)), (O.AGG_COUNT, E.col(2)),
                        (O.AGG_COUNT_STAR, None), (O.AGG_SUM, E.col(1) * (E.f64(1.0) - E.col(2)))])
    assert res.rows_filtered == want.rows_filtered
    assert (None,) in res.by_key()  # NULL keys form one group
    U.assert_agg_equal(res, want)
    # integer sums are bit-exact (Int64 wrapping), grouped by a string key
    res = scan.pipeline().aggregate([3], [(AggFunc.SUM, [Factor.of(4)]), (AggFunc.COUNT, [Factor.of(4)])]).run()
    want = O.aggregate(table, None, [E.col(3)], [(O.AGG_SUM, E.col(4)), (O.AGG_COUNT, E.col(4))])
    U.assert_agg_equal(res, want, rel=0)
    scan.release()


def test_reference_smoke_values(ctx):
    """pg/extension/src/smoke_tests.rs:205-251 re-expressed on synthetic pages."""
    ids = np.arange(1, 50_001, dtype=np.int64)
    schema = [ColumnSpec(TypeTag.Int64)]
    scan = load(ctx, schema, AL.encode_pages(schema, [(ids, None)]))
    res = scan.pipeline().aggregate([], [(AggFunc.AVG, [Factor.of(0)])]).run()
    assert abs(res.aggs[0][0] - 25000.5) < 1e-3            # smoke_tests.rs:205-226
    scan.release()
    ids = np.arange(1, 5001, dtype=np.int64)
    schema = [ColumnSpec(TypeTag.Int64), ColumnSpec(TypeTag.Utf8View)]
    pages = AL.encode_pages(schema, [(ids, None), (AL.inline_views([b"payload"] * 5000), None)])
    scan = load(ctx, schema, pages)
    res = scan.pipeline().aggregate([], [(AggFunc.COUNT, [Factor.of(0)]), (AggFunc.SUM, [Factor.of(0)])]).run(pages=True)
    assert res.aggs[0] == (5000, 12502500)                  # smoke_tests.rs:228-251
    # transport schema: COUNT is NOT NULL, SUM(Int64) is a nullable Int64 although this result holds no NULL
    assert [(int(c.type_tag), bool(c.nullable)) for c in res.result_schema] == [(4, False), (4, True)]
    res = scan.pipeline().filter(0, Cmp.GE, 10).filter(0, Cmp.LE, 20).aggregate([], [(AggFunc.SUM, [Factor.of(0)])]).run()
    assert res.aggs[0] == (sum(range(10, 21)),)             # smoke_tests.rs:382-471 (BETWEEN)
    scan.release()


def test_many_groups_take_the_hash_path(ctx):
    r = np.random.default_rng(3)
    n = 200_000
    schema = [ColumnSpec(TypeTag.Int32), ColumnSpec(TypeTag.Int64), ColumnSpec(TypeTag.Float64)]
    k = r.integers(0, 30_000, n).astype(np.int32)
    v = r.integers(-10**9, 10**9, n)
    f = r.random(n) * 1000
    pages = AL.encode_pages(schema, [(k, None), (v, None), (f, None)])
    scan = load(ctx, schema, pages)
    table = O.OTable.from_pages(pages, 65536, U.orc_cols(schema))
    res = scan.pipeline().aggregate([0], [(AggFunc.SUM, [Factor.of(1)]), (AggFunc.COUNT_STAR, None)], expected_groups=100).run()
    want = O.aggregate(table, None, [E.col(0)], [(O.AGG_SUM, E.col(1)), (O.AGG_COUNT_STAR, None)])
    assert len(res.keys) == len(want.keys)
    U.assert_agg_equal(res, want, rel=0)   # forces a table-overflow re-run (expected_groups far too low)
    res = scan.pipeline().aggregate([0], [(AggFunc.SUM, [Factor.of(2)]), (AggFunc.AVG, [Factor.of(2)])]).run()
    want = O.aggregate(table, None, [E.col(0)], [(O.AGG_SUM, E.col(2)), (O.AGG_AVG, E.col(2))])
    U.assert_agg_equal(res, want)
    scan.release()


def test_decimal128_extension_is_bit_exact(ctx):
    """Decimal128(15,2) money (SURVEY 8d "D" schema): i128 sums and products are bit-exact,
    AVG follows DecimalAverager (sum * 10^4 / count, truncating)."""
    r = np.random.default_rng(8)
    n = 30_000
    schema = [ColumnSpec(TypeTag.Decimal128), ColumnSpec(TypeTag.Decimal128), ColumnSpec(TypeTag.Decimal128), ColumnSpec(TypeTag.Int16)]

    def dec(vals):
        out = np.zeros((len(vals), 16), dtype=np.uint8)
        for i, v in enumerate(vals):
            out[i] = np.frombuffer((int(v) & (2**128 - 1)).to_bytes(16, "little"), dtype=np.uint8)
        return out
    price = r.integers(-10**13, 10**13, n)
    disc = r.integers(0, 11, n)
    tax = r.integers(0, 9, n)
    grp = r.integers(0, 3, n).astype(np.int16)
    pages = AL.encode_pages(schema, [(dec(price), None), (dec(disc), None), (dec(tax), None), (grp, None)])
    scan = load(ctx, schema, pages)
    table = O.OTable.from_pages(pages, 65536, U.orc_cols(schema))
    dp = [Factor.of(0), Factor.const_minus(100, 1)]
    res = (scan.pipeline().filter(1, Cmp.GE, 2).aggregate([3], [
        (AggFunc.SUM, [Factor.of(0)]), (AggFunc.SUM, dp), (AggFunc.SUM, dp + [Factor.const_plus(100, 2)]),
        (AggFunc.AVG, [Factor.of(0)]), (AggFunc.COUNT_STAR, None)]).run())
    d_price = E.col(0) * (E.i128(100) - E.col(1))
    want = O.aggregate(table, E.col(1).ge(E.i128(2)), [E.col(3)], [
        (O.AGG_SUM, E.col(0)), (O.AGG_SUM, d_price), (O.AGG_SUM, d_price * (E.i128(100) + E.col(2))),
        (O.AGG_SUM, E.col(0)), (O.AGG_COUNT_STAR, None)])
    got, exp = res.by_key(), want.by_key()
    assert set(got) == set(exp)
    for k in exp:
        s0, s1, s2, ssum, cnt = exp[k]
        q = abs(ssum * 10000) // cnt
        avg = q if ssum >= 0 else -q  # i128 division truncates toward zero
        assert got[k] == (s0, s1, s2, avg, cnt)
    scan.release()


def test_eligibility_and_unsupported_data(ctx):
    schema = [ColumnSpec(TypeTag.Boolean), ColumnSpec(TypeTag.Float64), ColumnSpec(TypeTag.Int32), ColumnSpec(TypeTag.Utf8View)]
    n = 100
    long_s = [b"a long string beyond twelve bytes"] * n
    blk = O.Block(U.orc_cols(schema), n, 65516)
    import struct
    for r_ in range(n):
        blk.write_bool(0, r_, True); blk.write_fixed(1, r_, struct.pack("<d", 1.0)); blk.write_fixed(2, r_, struct.pack("<i", r_))
        assert blk.write_view_bytes(3, r_, long_s[r_]) == 0
        blk.commit_current_row()
    page = np.concatenate([np.frombuffer(AL.page_header(), dtype=np.uint8), blk.buf])
    scan = load(ctx, schema, page.reshape(1, -1))
    # out-of-line views pass ingest (they are valid pages) ...
    assert scan.info().rows == n
    # ... plans over unsupported shapes are declined, like install_runtime_filters declines joins
    assert scan.pipeline().filter(0, Cmp.EQ, 1).count().check() == pg.errors.NOT_ELIGIBLE
    assert scan.pipeline().aggregate([], [(AggFunc.SUM, [Factor.of(1), Factor.of(2)])]).check() == pg.errors.NOT_ELIGIBLE
    assert scan.pipeline().filter(3, Cmp.EQ, b"x" * 13).count().check() == pg.errors.NOT_ELIGIBLE
    # predicates compare out-of-line values through the tail arena (test_out_of_line_view_values_in_predicates) ...
    assert scan.pipeline().filter(3, Cmp.EQ, b"a").count().run().rows_out == 0
    assert scan.pipeline().filter(3, Cmp.GT, b"a").count().run().rows_out == n
    # ... and data the kernel cannot use -- an out-of-line string as a GROUP BY key -- fails loudly instead of
    # returning a wrong answer
    with pytest.raises(pg.PgfError) as e:
        scan.pipeline().aggregate([3], [(AggFunc.COUNT_STAR, None)]).run()
    assert e.value.name == "UNSUPPORTED_DATA"
    assert scan.pipeline().filter(2, Cmp.LT, 10).count().run().rows_out == 10
    scan.release()


def test_ingest_rejects_bad_pages_on_host_and_device(ctx):
    schema = [ColumnSpec(TypeTag.Int32, True), ColumnSpec(TypeTag.Utf8View, True)]
    vals = np.arange(10, dtype=np.int32)
    valid = np.array([True] * 9 + [False])
    good = AL.encode_pages(schema, [(vals, valid), (AL.inline_views([b"ab"] * 10), valid)])
    # host-side structural rejection: wrong kind
    bad = good.copy(); bad[0, 10:12] = [0, 9]
    scan = ctx.declare_scan(schema)
    with pytest.raises(pg.PgfError) as e:
        scan.push_pages(bad)
    assert e.value.name == "IMPORT_WRONG_KIND"
    with pytest.raises(pg.PgfError) as e:
        scan.push_pages(good[:, :30000], stride=30000)
    assert e.value.name == "LAYOUT_BLOCK_SLICE_TOO_SMALL"
    # device-side row-level rejection: null bitmap popcount mismatch, invalid UTF-8
    plan = AL.LayoutPlan(schema, AL.fixed_row_cap(schema), 65516)
    for mutate, name in ((lambda p: p.__setitem__((0, 20 + plan.column_layout(0).validity_off + 1), 0xFF), "IMPORT_NULL_BITMAP_COUNT_MISMATCH"),
                         (lambda p: p.__setitem__((0, 20 + plan.column_layout(1).values_off + 4), 0xFF), "IMPORT_ARROW_INVALID_VIEW")):
        bad = good.copy(); mutate(bad)
        s2 = ctx.declare_scan(schema)
        s2.push_pages(bad)
        with pytest.raises(pg.PgfError) as e:
            s2.finish()
        assert e.value.name == name
        s2.release()
    scan.push_pages(good)
    scan.finish()
    assert scan.pipeline().count().run().rows_out == 10
    scan.release()


def test_q1_result_pages_round_trip(ctx):
    """R1: the aggregate output leaves as reference result pages (ResultPageProducer format) that
    the import path (oracle restatement of ArrowPageDecoder::import_owned) decodes to the same rows."""
    scan = ctx.gen_scan(pg.GenTable.LINEITEM_Q1, 50_000, seed=42)
    res = U.gpu_q1(scan).run(pages=True)
    cols = [(int(c.type_tag), bool(c.nullable)) for c in res.result_schema]
    assert [c[0] for c in cols] == [8, 8] + [6] * 7 + [4]      # 2 Utf8View keys, 7 Float64, COUNT(*) Int64
    # l_returnflag / l_linestatus are NOT NULL columns: their group fields are non-nullable; SUM / AVG nullable; COUNT(*) not
    assert [c[1] for c in cols] == [False, False] + [True] * 7 + [False]
    t = O.OTable.from_pages(res.result_pages, 65536, cols)
    assert t.rows == len(res.keys) == 4
    decoded = {}
    for g in range(t.rows):
        key = (t.column(0)[g], t.column(1)[g])
        decoded[key] = tuple(float(t.column(2 + j)[0][g]) for j in range(7)) + (int(t.column(9)[0][g]),)
    assert decoded == {k: tuple(v) for k, v in res.by_key().items()}
    scan.release()


def test_q1_shape_with_non_finite_values(ctx):
    """Inf / NaN arguments stay inside their own group and aggregate (IEEE semantics per group, as
    DataFusion's per-group accumulators): no selector trick in the sink may leak them elsewhere."""
    import math
    li = U.lineitem(30_000, 5)
    li["price"][::1000] = np.inf
    li["qty"][7::1500] = -np.inf
    li["disc"][11::2000] = np.nan
    pages = U.q1_pages(li)
    scan = load(ctx, U.Q1_SCHEMA, pages)
    res = U.gpu_q1(scan).run()
    want = U.oracle_q1(O.OTable.from_pages(pages, 65536, U.orc_cols(U.Q1_SCHEMA)))
    gk, ok = res.by_key(), want.by_key()
    assert set(gk) == set(ok)
    for k in ok:
        for j, (x, y) in enumerate(zip(gk[k], ok[k])):
            if isinstance(y, float) and not math.isfinite(y):
                assert (math.isnan(x) and math.isnan(y)) or x == y, (k, j, x, y)
            else:
                U.assert_close(x, y, 1e-12, f"group {k} agg {j}")
    # a group without any non-finite input keeps finite sums
    assert any(all(math.isfinite(v) for v in vals if isinstance(v, float)) for vals in gk.values()) or len(gk) < 4
    scan.release()


def test_scan_with_several_layout_classes_and_ragged_pages(ctx):
    """Pages of one scan may differ in max_rows (each page is self-describing, raw.rs:21-46): the
    producer switches layout classes per page; the last page of each run is partially filled."""
    parts, pages = [], []
    for n, seed, rpp in ((5_000, 1, 700), (3_333, 2, 1614), (777, 3, 128), (4_100, 4, 1000)):
        li = U.lineitem(n, seed)
        parts.append(li)
        pages.append(U.q6_pages(li, rows_per_page=rpp))
    allp = np.concatenate(pages)
    scan = ctx.declare_scan(U.Q6_SCHEMA)
    for p in pages:            # four pushes, interleaving the classes would work the same
        scan.push_pages(p)
    scan.finish()
    res = U.gpu_q6(scan).run()
    want = U.oracle_q6(O.OTable.from_pages(allp, 65536, U.orc_cols(U.Q6_SCHEMA)))
    assert res.rows_in == sum(len(p["qty"]) for p in parts) and res.rows_filtered == want.rows_filtered
    U.assert_agg_equal(res, want)
    # the same pages in an interleaved order (class changes on almost every page)
    order = np.random.default_rng(0).permutation(allp.shape[0])
    scan2 = ctx.declare_scan(U.Q6_SCHEMA)
    scan2.push_pages(np.ascontiguousarray(allp[order]))
    scan2.finish()
    res2 = U.gpu_q6(scan2).run()
    assert res2.rows_filtered == want.rows_filtered and res2.aggs[0][1] == want.aggs[0][1]
    U.assert_close(res2.aggs[0][0], want.aggs[0][0], 1e-12, "revenue, interleaved classes")
    scan.release()
    scan2.release()


def test_decimal_fast_group_by_is_exact_beyond_64_bits(ctx):
    """The registered Decimal128 Q1 shape keeps signed 64-bit partial sums in shared memory: values
    that do not fit 64 bits and additions that would overflow a slot must reach the 128-bit global
    accumulators, so the result is the exact wrapping i128 sum whatever the magnitudes."""
    r = np.random.default_rng(21)
    n = 40_000

    def dec(vals):
        out = np.zeros((len(vals), 16), dtype=np.uint8)
        for i, v in enumerate(vals):
            out[i] = np.frombuffer((int(v) & (2**128 - 1)).to_bytes(16, "little"), dtype=np.uint8)
        return out
    qty = [int(x) for x in r.integers(1, 51, n) * 100]
    price = [int(x) for x in r.integers(-10**13, 10**13, n)]
    for i in range(0, n, 7):        # slots overflow after a few rows
        price[i] = int(r.integers(2**62, 2**63 - 1)) * (1 if i % 14 else -1)
    for i in range(3, n, 11):       # values beyond 64 bits
        price[i] = int(r.integers(-10**18, 10**18)) * 10**12
    disc = [int(x) for x in r.integers(0, 11, n)]
    tax = [int(x) for x in r.integers(0, 9, n)]
    rf = r.choice(np.array([65, 78, 82], dtype=np.int16), n)
    ls = r.choice(np.array([70, 79], dtype=np.int16), n)
    ship = r.integers(8036, 10562, n).astype(np.int32)
    pages = AL.encode_pages(U.Q1_D_SCHEMA, [(dec(qty), None), (dec(price), None), (dec(disc), None), (dec(tax), None),
                                            (rf, None), (ls, None), (ship, None)])
    scan = load(ctx, U.Q1_D_SCHEMA, pages)
    res = U.gpu_q1_d(scan).run()
    want, raw = U.oracle_q1_d(O.OTable.from_pages(pages, 65536, U.orc_cols(U.Q1_D_SCHEMA)))
    assert res.rows_filtered == raw.rows_filtered and len(want) == 6
    wrap = lambda v: ((v + 2**127) % 2**128) - 2**127    # wrapping i128 (DataFusion's Decimal128 sum wraps)
    got = {k: tuple(v) for k, v in res.by_key().items()}
    assert set(got) == set(want)
    for k, w in want.items():
        sums = tuple(wrap(x) for x in w[:4])
        assert got[k][:4] == sums and got[k][7] == w[7], k
    scan.release()


def test_registered_shapes_match_whatever_the_order_of_conjuncts_and_aggregates(ctx):
    """Conjunct order and aggregate order are irrelevant to the result; they must not decide whether
    the plan gets a specialised kernel either (pgf_result.variant names the instantiation)."""
    import itertools
    scan = ctx.gen_scan(pg.GenTable.LINEITEM_Q6, 200_000, seed=42)
    base = U.gpu_q6(scan).run()
    assert base.variant == "q6_f64"
    conj = [(3, Cmp.GE, b"1994-01-01"), (3, Cmp.LT, b"1995-01-01"), (2, Cmp.GE, 0.05), (2, Cmp.LE, 0.07), (0, Cmp.LT, 24.0)]
    for perm in list(itertools.permutations(conj))[::17]:
        p = scan.pipeline()
        for c in perm:
            p.filter(*c)
        r = p.aggregate([], [(AggFunc.COUNT_STAR, None), (AggFunc.SUM, [Factor.of(1), Factor.of(2)])]).run()
        assert r.variant == "q6_f64"
        assert r.aggs[0][0] == base.aggs[0][1] and r.rows_filtered == base.rows_filtered
        U.assert_close(r.aggs[0][1], base.aggs[0][0], 1e-12, "revenue")
    scan.release()
    q1 = ctx.gen_scan(pg.GenTable.LINEITEM_Q1, 200_000, seed=42)
    ref = U.gpu_q1(q1).run()
    assert ref.variant == "q1_f64_8aggs"
    q, pr, d, t, rf, ls, s = range(7)
    disc_price = [Factor.of(pr), Factor.const_minus(1.0, d)]
    charge = disc_price + [Factor.const_plus(1.0, t)]
    aggs = [(AggFunc.AVG, [Factor.of(d)]), (AggFunc.SUM, charge), (AggFunc.COUNT_STAR, None), (AggFunc.SUM, [Factor.of(pr)]),
            (AggFunc.AVG, [Factor.of(q)]), (AggFunc.SUM, disc_price), (AggFunc.SUM, [Factor.of(q)]), (AggFunc.AVG, [Factor.of(pr)])]
    got = q1.pipeline().filter(s, Cmp.LE, b"1998-09-02").aggregate([rf, ls], aggs).order_by([("agg", 1, True)]).run()
    assert got.variant == "q1_f64_8aggs"
    back = [6, 3, 5, 1, 4, 7, 0, 2]   # position in `aggs` of each aggregate of U.gpu_q1's order
    want = ref.by_key()
    assert [a[1] for a in got.aggs] == sorted((a[1] for a in got.aggs), reverse=True)   # ORDER BY sum_charge DESC
    for k, a in zip(got.keys, got.aggs):
        for j in range(8):
            U.assert_close(a[back[j]], want[k][j], 1e-12, f"group {k} agg {j}")
    q1.release()


def _pages_with_long_strings(n, seed, max_rows=400):
    """Pages written row by row with the oracle's BlockMut restatement: strings of 0..40 bytes over a two-letter alphabet
    (so that many share their first 4 and their first 12 bytes), the long ones out of line in the tail arena
    (page/arrow_layout/src/raw.rs:98-110).  Columns: Int32 id, nullable Utf8View s, Float64 v."""
    import struct
    r = np.random.default_rng(seed)
    cols = [(O.T_INT32, False), (O.T_UTF8VIEW, True), (O.T_FLOAT64, False)]
    pages, row = [], 0
    while row < n:
        blk = O.Block(cols, max_rows, 65516)
        m = min(max_rows, n - row)
        for i in range(m):
            blk.write_fixed(0, i, struct.pack("<i", row + i))
            ln = int(r.integers(0, 41))
            if r.random() < 0.08:
                blk.write_null(1, i)
            else:
                s = bytes(r.choice([65, 66], ln).astype(np.uint8)) if r.random() < 0.7 else b"ABABABABABAB"[:min(ln, 12)] + b"A" * max(0, ln - 12)
                assert blk.write_view_bytes(1, i, s) == 0
            blk.write_fixed(2, i, struct.pack("<d", float(r.integers(1, 1000)) / 8.0))
            blk.commit_current_row()
        page = np.zeros(65536, np.uint8)
        page[:20] = np.frombuffer(O.page_header(0x4152, 0, 65516), np.uint8)
        page[20:] = blk.buf
        pages.append(page)
        row += m
    return np.stack(pages)


def test_out_of_line_view_values_in_predicates(ctx):
    """VERDICT r1 grammar gap / SURVEY 8(f)4: string values longer than 12 bytes live in the page's tail arena; a
    predicate over them compares the first 12 bytes fetched from there and the length.  Both kernels (streaming
    aggregate, compaction pipeline behind a join) against the oracle, which compares the full byte strings."""
    n = 6000
    schema = [ColumnSpec(TypeTag.Int32), ColumnSpec(TypeTag.Utf8View, True), ColumnSpec(TypeTag.Float64)]
    pages = _pages_with_long_strings(n, 5)
    assert AL.import_check(0x4152, 0, pages[0][20:], schema) == 0
    scan = load(ctx, schema, pages)
    table = O.OTable.from_pages(pages, 65536, U.orc_cols(schema))
    strings = table.column(1)
    assert sum(1 for s in strings if s is not None and len(s) > 12) > n // 3
    # a build side for the compaction pipeline: every third id
    ids = np.arange(0, n, 3, dtype=np.int32)
    bs = [ColumnSpec(TypeTag.Int32)]
    build = load(ctx, bs, AL.encode_pages(bs, [(ids, None)]))
    jt = build.pipeline().build_join(0, []).run()
    agg = [(AggFunc.SUM, [Factor.of(2)]), (AggFunc.COUNT_STAR, None)]
    cases = [[(Cmp.GT, b"ABAB")], [(Cmp.LE, b"ABABABABABAB")], [(Cmp.EQ, b"ABABABABABAB")], [(Cmp.NE, b"ABABABABABAB")],
             [(Cmp.GE, b"AB"), (Cmp.LT, b"ABB")], [(Cmp.GT, b"ABABABABABAB"), (Cmp.LT, b"B")], [(Cmp.LT, b"")], [(Cmp.GE, b"BBBBBBBBBBBB")]]
    ops = {Cmp.GT: "gt", Cmp.GE: "ge", Cmp.LT: "lt", Cmp.LE: "le", Cmp.EQ: "eq", Cmp.NE: "ne"}
    for terms in cases:
        p = scan.pipeline()
        filt = None
        for cmp, lit in terms:
            p = p.filter(1, cmp, lit)
            t = getattr(E.col(1), ops[cmp])(E.s(lit))
            filt = t if filt is None else filt.and_(t)
        res = p.aggregate([], agg).run()
        want = O.aggregate(table, filt, [], [(O.AGG_SUM, E.col(2)), (O.AGG_COUNT_STAR, None)])
        # ground truth straight from the decoded strings (Python compares bytes the way arrow does)
        py = {Cmp.GT: lambda a, b: a > b, Cmp.GE: lambda a, b: a >= b, Cmp.LT: lambda a, b: a < b, Cmp.LE: lambda a, b: a <= b,
              Cmp.EQ: lambda a, b: a == b, Cmp.NE: lambda a, b: a != b}
        kept = sum(1 for s in strings if s is not None and all(py[c](s, l) for c, l in terms))
        assert res.rows_filtered == want.rows_filtered == kept, terms
        U.assert_agg_equal(res, want)
        # the same predicate in front of a join probe (compaction pipeline)
        pj = scan.pipeline()
        for cmp, lit in terms:
            pj = pj.filter(1, cmp, lit)
        rj = pj.join(jt.join_table, 0).aggregate([], agg).run()
        keptj = sum(1 for i, s in enumerate(strings) if i % 3 == 0 and s is not None and all(py[c](s, l) for c, l in terms))
        assert rj.rows_filtered == kept and rj.aggs[0][1] == keptj, terms
    scan.release()
    build.release()


@pytest.mark.parametrize("rows_per_page", [None, 300])
def test_boolean_columns_in_predicates(ctx, rows_per_page):
    """VERDICT r1 grammar gap: Boolean columns are bit-packed on the page (page/arrow_layout/src/types.rs:139-147); a
    predicate over one stages the bitmap slice of the tile.  WHERE flag [= true] / flag = false / flag <> true, alone
    and with other conjuncts, NULL => dropped; streaming kernel and compaction pipeline."""
    r = np.random.default_rng(9)
    n = 30_000
    schema = [ColumnSpec(TypeTag.Boolean, True), ColumnSpec(TypeTag.Int64), ColumnSpec(TypeTag.Boolean), ColumnSpec(TypeTag.Float64)]
    f0, f2 = r.random(n) < 0.4, r.random(n) < 0.7
    v0 = r.random(n) > 0.1
    k = r.integers(0, 5000, n)
    x = r.integers(1, 1000, n) / 4.0
    pages = AL.encode_pages(schema, [(f0, v0), (k, None), (f2, None), (x, None)], rows_per_page=rows_per_page)
    assert O.import_check(0x4152, 0, np.ascontiguousarray(pages[0][20:]), U.orc_cols(schema)) == 0
    scan = load(ctx, schema, pages)
    table = O.OTable.from_pages(pages, 65536, U.orc_cols(schema))
    agg = [(AggFunc.SUM, [Factor.of(3)]), (AggFunc.COUNT_STAR, None)]
    oagg = [(O.AGG_SUM, E.col(3)), (O.AGG_COUNT_STAR, None)]
    # WHERE f0 (nullable)  -- the oracle evaluates the Boolean column itself as the predicate
    res = scan.pipeline().filter(0, Cmp.EQ, True).aggregate([], agg).run()
    want = O.aggregate(table, E.col(0), [], oagg)
    assert res.rows_filtered == want.rows_filtered == int((f0 & v0).sum())
    U.assert_agg_equal(res, want)
    # WHERE f0 = false: NULL is not FALSE
    res = scan.pipeline().filter(0, Cmp.EQ, False).aggregate([], agg).run()
    assert res.rows_filtered == int((~f0 & v0).sum()) and res.aggs[0][1] == res.rows_filtered
    U.assert_close(res.aggs[0][0], float(x[~f0 & v0].sum()), 1e-12, "sum over f0 = false")
    # WHERE f0 <> true AND f2 AND k < 2500, grouped
    res = (scan.pipeline().filter(0, Cmp.NE, True).filter(2, Cmp.EQ, True).filter(1, Cmp.LT, 2500)
           .aggregate([1], [(AggFunc.SUM, [Factor.of(3)]), (AggFunc.COUNT_STAR, None)]).run())
    m = ~f0 & v0 & f2 & (k < 2500)
    assert res.rows_filtered == int(m.sum())
    want = O.aggregate(table, E.col(2).and_(E.col(1).lt(E.i64(2500))).and_(E.col(0).eq(E.col(0)).and_(E.col(0).ne(E.col(2).eq(E.col(2))))), [E.col(1)], oagg)
    assert want.rows_filtered == int(m.sum())   # (f0 <> true written as f0 <> (f2 = f2) for the oracle's expression grammar)
    U.assert_agg_equal(res, want)
    # the compaction pipeline: the same Boolean predicate in front of a join probe
    bs = [ColumnSpec(TypeTag.Int64)]
    build = load(ctx, bs, AL.encode_pages(bs, [(np.arange(0, 5000, 2, dtype=np.int64), None)]))
    jt = build.pipeline().build_join(0, []).run()
    rj = scan.pipeline().filter(2, Cmp.EQ, True).filter(0, Cmp.EQ, False).join(jt.join_table, 1).aggregate([], agg).run()
    mj = f2 & ~f0 & v0
    assert rj.rows_filtered == int(mj.sum()) and rj.aggs[0][1] == int((mj & (k % 2 == 0)).sum())
    U.assert_close(rj.aggs[0][0], float(x[mj & (k % 2 == 0)].sum()), 1e-12, "join behind a Boolean predicate")
    # a Boolean column is not a GROUP BY key or an aggregate argument on this path
    with pytest.raises(pg.PgfError) as e:
        scan.pipeline().aggregate([0], agg).run()
    assert e.value.code == 6   # PGF_ERR_NOT_ELIGIBLE: the caller keeps the DataFusion node
    scan.release()
    build.release()
