"""Shared helpers for the parity tests: seeded synthetic tables, page fabrication through the
product's host writer, and the oracle-side formulation of the TPC-H shapes."""
from __future__ import annotations

import datetime as dt
from typing import List, Optional, Sequence, Tuple

import numpy as np

from oracle import pyorc as O
from pg_fusion_b200 import ColumnSpec, TypeTag
from pg_fusion_b200 import arrow_layout as AL

F64, I32, I64, I16, VIEW = TypeTag.Float64, TypeTag.Int32, TypeTag.Int64, TypeTag.Int16, TypeTag.Utf8View

Q6_SCHEMA = [ColumnSpec(F64), ColumnSpec(F64), ColumnSpec(F64), ColumnSpec(VIEW)]
Q1_SCHEMA = [ColumnSpec(F64)] * 4 + [ColumnSpec(VIEW)] * 3


def orc_cols(schema: Sequence[ColumnSpec]):
    return [(int(c.type_tag), bool(c.nullable)) for c in schema]


def dates_from_days(days: np.ndarray) -> List[bytes]:
    base = dt.date(1992, 1, 1)
    return [(base + dt.timedelta(days=int(d))).isoformat().encode() for d in days]


def lineitem(n: int, seed: int):
    """TPC-H-shaped lineitem columns: money as f64 = cents / 100.0."""
    r = np.random.default_rng(seed)
    qty = r.integers(1, 51, n)
    part = r.integers(90000, 210001, n)
    price = (qty * part) / 100.0
    disc = r.integers(0, 11, n) / 100.0
    tax = r.integers(0, 9, n) / 100.0
    ship = r.integers(1, 2527, n)
    receipt = ship + r.integers(1, 31, n)
    cutoff = 1263  # 1995-06-17
    ls = np.where(ship > cutoff, b"O", b"F")
    rf = np.where(receipt <= cutoff, np.where(r.integers(0, 2, n) == 1, b"R", b"A"), b"N")
    return dict(qty=qty.astype(np.float64), price=price, disc=disc, tax=tax, ship=ship, rf=rf, ls=ls)


def q6_pages(li, page_size=65536, rows_per_page=None) -> np.ndarray:
    cols = [(li["qty"], None), (li["price"], None), (li["disc"], None),
            (AL.inline_views(dates_from_days(li["ship"])), None)]
    return AL.encode_pages(Q6_SCHEMA, cols, page_size, rows_per_page)


def q1_pages(li, page_size=65536, rows_per_page=None) -> np.ndarray:
    cols = [(li["qty"], None), (li["price"], None), (li["disc"], None), (li["tax"], None),
            (AL.inline_views([bytes(x) for x in li["rf"]]), None),
            (AL.inline_views([bytes(x) for x in li["ls"]]), None),
            (AL.inline_views(dates_from_days(li["ship"])), None)]
    return AL.encode_pages(Q1_SCHEMA, cols, page_size, rows_per_page)


E = O.Expr


def oracle_q6(table: O.OTable, cols=(0, 1, 2, 3), sum_lanes=0) -> O.AggOut:
    q, p, d, s = cols
    filt = (E.col(s).ge(E.s(b"1994-01-01"))).and_(E.col(s).lt(E.s(b"1995-01-01"))) \
        .and_(E.col(d).ge(E.f64(0.05))).and_(E.col(d).le(E.f64(0.07))).and_(E.col(q).lt(E.f64(24.0)))
    return O.aggregate(table, filt, [], [(O.AGG_SUM, E.col(p) * E.col(d)), (O.AGG_COUNT_STAR, None)], sum_lanes=sum_lanes)


def oracle_q1(table: O.OTable) -> O.AggOut:
    q, p, d, t, rf, ls, s = range(7)
    filt = E.col(s).le(E.s(b"1998-09-02"))
    disc_price = E.col(p) * (E.f64(1.0) - E.col(d))
    charge = disc_price * (E.f64(1.0) + E.col(t))
    aggs = [(O.AGG_SUM, E.col(q)), (O.AGG_SUM, E.col(p)), (O.AGG_SUM, disc_price), (O.AGG_SUM, charge),
            (O.AGG_AVG, E.col(q)), (O.AGG_AVG, E.col(p)), (O.AGG_AVG, E.col(d)), (O.AGG_COUNT_STAR, None)]
    return O.aggregate(table, filt, [E.col(rf), E.col(ls)], aggs)


def gpu_q6(scan, cols=(0, 1, 2, 3)):
    from pg_fusion_b200 import AggFunc, Cmp, Factor
    q, p, d, s = cols
    return (scan.pipeline()
            .filter(s, Cmp.GE, b"1994-01-01").filter(s, Cmp.LT, b"1995-01-01")
            .filter(d, Cmp.GE, 0.05).filter(d, Cmp.LE, 0.07).filter(q, Cmp.LT, 24.0)
            .aggregate([], [(AggFunc.SUM, [Factor.of(p), Factor.of(d)]), (AggFunc.COUNT_STAR, None)]))


def gpu_q1(scan):
    from pg_fusion_b200 import AggFunc, Cmp, Factor
    q, p, d, t, rf, ls, s = range(7)
    disc_price = [Factor.of(p), Factor.const_minus(1.0, d)]
    charge = disc_price + [Factor.const_plus(1.0, t)]
    aggs = [(AggFunc.SUM, [Factor.of(q)]), (AggFunc.SUM, [Factor.of(p)]), (AggFunc.SUM, disc_price),
            (AggFunc.SUM, charge), (AggFunc.AVG, [Factor.of(q)]), (AggFunc.AVG, [Factor.of(p)]),
            (AggFunc.AVG, [Factor.of(d)]), (AggFunc.COUNT_STAR, None)]
    return scan.pipeline().filter(s, Cmp.LE, b"1998-09-02").aggregate([rf, ls], aggs)


def assert_close(a, b, rel=1e-12, what=""):
    """Float64 SUM/AVG tolerance stated by BASELINE.json's north_star: 1e-12 relative."""
    if a is None or b is None:
        assert a is None and b is None, f"{what}: {a} vs {b}"
        return
    if isinstance(a, float) or isinstance(b, float):
        denom = max(abs(a), abs(b), 1e-300)
        assert abs(a - b) / denom <= rel, f"{what}: {a!r} vs {b!r} rel={abs(a - b) / denom:.3e}"
    else:
        assert a == b, f"{what}: {a!r} vs {b!r}"


def assert_agg_equal(gpu, orc, rel=1e-12):
    gk, ok = gpu.by_key(), orc.by_key()
    assert set(gk) == set(ok), f"group keys differ: {sorted(gk)[:5]} vs {sorted(ok)[:5]}"
    for k in ok:
        assert len(gk[k]) == len(ok[k])
        for j, (x, y) in enumerate(zip(gk[k], ok[k])):
            assert_close(x, y, rel, f"group {k} agg {j}")


# ---- TPC-H Q3 shape: customer |><| orders |><| lineitem, benches/tpch/queries/q03.sql ----
CUSTOMER_SCHEMA = [ColumnSpec(I32), ColumnSpec(VIEW)]                                    # c_custkey, c_mktsegment
ORDERS_SCHEMA = [ColumnSpec(I32), ColumnSpec(I32), ColumnSpec(VIEW), ColumnSpec(I32)]    # o_orderkey, o_custkey, o_orderdate, o_shippriority
LINEITEM_Q3_SCHEMA = [ColumnSpec(I32), ColumnSpec(F64), ColumnSpec(F64), ColumnSpec(VIEW)]  # l_orderkey, l_extendedprice, l_discount, l_shipdate
Q3_DATE = b"1995-03-15"


Q3_ORDER = [("agg", 0, True), ("key", 1, False)]   # ORDER BY revenue DESC, o_orderdate (q03.sql)


def gpu_q3(ctx, customer, orders, lineitem, bloom_params=None, segment=b"BUILDING", limit=0):
    """Runs the three fused pipelines of the Q3 shape; returns (result, stats dict).
    limit > 0 adds ORDER BY revenue DESC, o_orderdate LIMIT n (device top-k)."""
    import pg_fusion_b200 as pg
    from pg_fusion_b200 import AggFunc, Cmp, Factor
    stats = {}
    rf1 = rf2 = None
    if bloom_params is not None:
        rf1 = ctx.runtime_filter(bloom_params[0])
        rf1.try_acquire_builder()
    # customer(BUILDING) -> join table T1 keyed by c_custkey (+ Bloom for the orders scan)
    r1 = customer.pipeline().filter(1, Cmp.EQ, segment).build_join(0, [], rf1).run()
    if rf1 is not None:
        rf1.publish_ready()
    # orders: [Bloom probe] -> o_orderdate < date -> probe T1 -> T2 keyed by o_orderkey with payload
    p2 = orders.pipeline()
    if rf1 is not None:
        p2.bloom_probe(rf1, 1)
        rf2 = ctx.runtime_filter(bloom_params[1])
        rf2.try_acquire_builder()
    r2 = p2.filter(2, Cmp.LT, Q3_DATE).join(r1.join_table, 1).build_join(0, [2, 3], rf2).run()
    if rf2 is not None:
        rf2.publish_ready()
    # lineitem: [Bloom probe] -> l_shipdate > date -> probe T2 -> GROUP BY l_orderkey, o_orderdate, o_shippriority
    p3 = lineitem.pipeline()
    if rf2 is not None:
        p3.bloom_probe(rf2, 0)
    p3 = (p3.filter(3, Cmp.GT, Q3_DATE).join(r2.join_table, 0)
          .aggregate([0, (1, 0), (1, 1)], [(AggFunc.SUM, [Factor.of(1), Factor.const_minus(1.0, 2)])],
                     expected_groups=max(1024, r2.rows_out)))
    if limit:
        p3.order_by(Q3_ORDER, limit=limit)
    r3 = p3.run()
    ctx.destroy_join_table(r1.join_table)
    ctx.destroy_join_table(r2.join_table)
    stats.update(customer=r1, orders=r2, lineitem=r3, rf1=rf1, rf2=rf2)
    return r3, stats


def oracle_q3(customer_t, orders_t, lineitem_t, segment=b"BUILDING"):
    cust_f = customer_t.select(O.filter_rows(customer_t, E.col(1).eq(E.s(segment))))
    ord_f = orders_t.select(O.filter_rows(orders_t, E.col(2).lt(E.s(Q3_DATE))))
    _, probe_rows = O.hash_join_pairs(cust_f, 0, ord_f, 1)      # HashJoinExec(customer, orders)
    ord_j = ord_f.take(probe_rows)
    res = O.aggregate(lineitem_t, E.col(3).gt(E.s(Q3_DATE)), [E.col(0), E.col(2, 1), E.col(3, 1)],
                      [(O.AGG_SUM, E.col(1) * (E.f64(1.0) - E.col(2)))], joins=[(ord_j, 0, 0, 0)])
    return res, dict(customers=cust_f.rows, orders=ord_j.rows)


def top10(res):
    """ORDER BY revenue DESC, o_orderdate LIMIT 10 (benches/tpch/queries/q03.sql)."""
    rows = [(k[0], a[0], k[1], k[2]) for k, a in zip(res.keys, res.aggs)]
    return sorted(rows, key=lambda r: (-r[1], r[2]))[:10]


def gpu_q3_sharded(ctx, customer, orders, lineitem, world, device, bloom_params=None, segment=b"BUILDING", limit=0):
    """The Q3 shape with every scan sharded by pages over `world` ranks (one process per GPU):
    broadcast joins, OR-merged runtime filters, Partial -> Final aggregate."""
    from pg_fusion_b200 import AggFunc, Cmp, Factor
    from pg_fusion_b200 import multi_gpu as MG
    rf1 = rf2 = None
    if bloom_params is not None:
        rf1 = ctx.runtime_filter(bloom_params[0])
        rf1.try_acquire_builder()
    r1 = customer.pipeline().filter(1, Cmp.EQ, segment).build_join(0, [], rf1).run()
    t1 = MG.broadcast_join_table(ctx, r1.join_table, world, device)
    if rf1 is not None:
        MG.or_merge_filter(rf1, world, device)
        rf1.publish_ready()
    p2 = orders.pipeline()
    if rf1 is not None:
        p2.bloom_probe(rf1, 1)
        rf2 = ctx.runtime_filter(bloom_params[1])
        rf2.try_acquire_builder()
    r2 = p2.filter(2, Cmp.LT, Q3_DATE).join(t1, 1).build_join(0, [2, 3], rf2).run()
    t2 = MG.broadcast_join_table(ctx, r2.join_table, world, device)
    if rf2 is not None:
        MG.or_merge_filter(rf2, world, device)
        rf2.publish_ready()
    p3 = lineitem.pipeline()
    if rf2 is not None:
        p3.bloom_probe(rf2, 0)
    total_orders = ctx.join_table_info(t2).rows
    p3 = (p3.filter(3, Cmp.GT, Q3_DATE).join(t2, 0)
          .aggregate([0, (1, 0), (1, 1)], [(AggFunc.SUM, [Factor.of(1), Factor.const_minus(1.0, 2)])],
                     expected_groups=max(1024, total_orders)))
    if limit:
        p3.order_by(Q3_ORDER, limit=limit)   # applied to the merged (final) groups
    res, stats = MG.merge_partial_aggregate(p3, world, device, max_groups=max(1024, total_orders))
    ctx.destroy_join_table(t1)
    ctx.destroy_join_table(t2)
    return res, dict(customer=r1, orders=r2, lineitem=stats)


# ---- "D" schema variants: money Decimal128(15,2) as unscaled hundredths in 16-byte slots,
# dates Date32 (Int32 days since 1970-01-01), flags Int16 character codes.  Decimal arithmetic follows
# DataFusion's type rules: 1 - disc -> (100 - disc) at scale 2, products are plain i128 products.
DEC = TypeTag.Decimal128
Q6_D_SCHEMA = [ColumnSpec(DEC), ColumnSpec(DEC), ColumnSpec(DEC), ColumnSpec(I32)]
Q1_D_SCHEMA = [ColumnSpec(DEC)] * 4 + [ColumnSpec(I16), ColumnSpec(I16), ColumnSpec(I32)]
D_1994, D_1995, D_1998_09_02 = 8766, 9131, 10471   # days since 1970-01-01


def gpu_q6_d(scan):
    from pg_fusion_b200 import AggFunc, Cmp, Factor
    return (scan.pipeline().filter(3, Cmp.GE, D_1994).filter(3, Cmp.LT, D_1995)
            .filter(2, Cmp.GE, 5).filter(2, Cmp.LE, 7).filter(0, Cmp.LT, 2400)
            .aggregate([], [(AggFunc.SUM, [Factor.of(1), Factor.of(2)]), (AggFunc.COUNT_STAR, None)]))


def oracle_q6_d(table: O.OTable) -> O.AggOut:
    filt = (E.col(3).ge(E.i64(D_1994))).and_(E.col(3).lt(E.i64(D_1995))) \
        .and_(E.col(2).ge(E.i128(5))).and_(E.col(2).le(E.i128(7))).and_(E.col(0).lt(E.i128(2400)))
    return O.aggregate(table, filt, [], [(O.AGG_SUM, E.col(1) * E.col(2)), (O.AGG_COUNT_STAR, None)])


def gpu_q1_d(scan):
    from pg_fusion_b200 import AggFunc, Cmp, Factor
    disc_price = [Factor.of(1), Factor.const_minus(100, 2)]
    charge = disc_price + [Factor.const_plus(100, 3)]
    aggs = [(AggFunc.SUM, [Factor.of(0)]), (AggFunc.SUM, [Factor.of(1)]), (AggFunc.SUM, disc_price), (AggFunc.SUM, charge),
            (AggFunc.AVG, [Factor.of(0)]), (AggFunc.AVG, [Factor.of(1)]), (AggFunc.AVG, [Factor.of(2)]), (AggFunc.COUNT_STAR, None)]
    return scan.pipeline().filter(6, Cmp.LE, D_1998_09_02).aggregate([4, 5], aggs)


def oracle_q1_d(table: O.OTable):
    """Returns {key: tuple} with DecimalAverager semantics for AVG (sum * 10^4 / count, truncating)."""
    disc_price = E.col(1) * (E.i128(100) - E.col(2))
    charge = disc_price * (E.i128(100) + E.col(3))
    res = O.aggregate(table, E.col(6).le(E.i64(D_1998_09_02)), [E.col(4), E.col(5)],
                      [(O.AGG_SUM, E.col(0)), (O.AGG_SUM, E.col(1)), (O.AGG_SUM, disc_price), (O.AGG_SUM, charge),
                       (O.AGG_SUM, E.col(2)), (O.AGG_COUNT_STAR, None)])
    out = {}
    for k, (sq, sp, sdp, sch, sd, cnt) in res.by_key().items():
        avg = lambda s: (abs(s * 10000) // cnt) * (1 if s >= 0 else -1)
        out[k] = (sq, sp, sdp, sch, avg(sq), avg(sp), avg(sd), cnt)
    return out, res
