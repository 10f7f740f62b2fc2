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
from pg_fusion_b200.tpch import gpu_q1, gpu_q1_d, gpu_q3, gpu_q3_sharded, gpu_q6, gpu_q6_d  # noqa: E402,F401  (the plans live in the package)


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


# ---- "D" schema variants: money Decimal128(15,2) as unscaled hundredths in 16-byte slots,
# dates Date32 (Int32 days since 1970-01-01), flags Int16 character codes.  Decimal arithmetic follows
# DataFusion's type rules: 1 - disc -> (100 - disc) at scale 2, products are plain i128 products.
DEC = TypeTag.Decimal128
Q6_D_SCHEMA = [ColumnSpec(DEC), ColumnSpec(DEC), ColumnSpec(DEC), ColumnSpec(I32)]
Q1_D_SCHEMA = [ColumnSpec(DEC)] * 4 + [ColumnSpec(I16), ColumnSpec(I16), ColumnSpec(I32)]
D_1994, D_1995, D_1998_09_02 = 8766, 9131, 10471   # days since 1970-01-01


def oracle_q6_d(table: O.OTable) -> O.AggOut:
    filt = (E.col(3).ge(E.i64(D_1994))).and_(E.col(3).lt(E.i64(D_1995))) \
        .and_(E.col(2).ge(E.i128(5))).and_(E.col(2).le(E.i128(7))).and_(E.col(0).lt(E.i128(2400)))
    return O.aggregate(table, filt, [], [(O.AGG_SUM, E.col(1) * E.col(2)), (O.AGG_COUNT_STAR, None)])


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


def q3_host_tables(ncust: int, nord: int, nli: int, seed: int, dup_keys: bool = False, rows_per_page=None):
    """Small Q3-shaped tables fabricated on the host (numpy): pages + oracle tables.  dup_keys: customer and
    order keys repeat, so both joins multiply rows (HashJoinExec keeps every pair)."""
    r = np.random.default_rng(seed)
    segs = [b"AUTOMOBILE", b"BUILDING", b"FURNITURE", b"MACHINERY", b"HOUSEHOLD"]
    ckey = (r.integers(1, max(2, ncust // 2), ncust) if dup_keys else np.arange(1, ncust + 1)).astype(np.int32)
    cseg = [segs[i] for i in r.integers(0, 5, ncust)]
    okey = (r.integers(1, max(2, nord // 2), nord) if dup_keys else np.arange(1, nord + 1) * 4).astype(np.int32)
    ocust = r.integers(1, ncust + 1, nord).astype(np.int32)
    odate = dates_from_days(r.integers(1, 2400, nord))
    oprio = r.integers(0, 3, nord).astype(np.int32)
    lkey = okey[r.integers(0, nord, nli)].astype(np.int32)
    lkey[r.random(nli) < 0.1] = -7                      # keys without a partner
    lprice = r.integers(90000, 10_000_000, nli) / 100.0
    ldisc = r.integers(0, 11, nli) / 100.0
    ldate = dates_from_days(r.integers(1, 2527, nli))
    pages = [AL.encode_pages(CUSTOMER_SCHEMA, [(ckey, None), (AL.inline_views(cseg), None)], 65536, rows_per_page),
             AL.encode_pages(ORDERS_SCHEMA, [(okey, None), (ocust, None), (AL.inline_views(odate), None), (oprio, None)], 65536, rows_per_page),
             AL.encode_pages(LINEITEM_Q3_SCHEMA, [(lkey, None), (lprice, None), (ldisc, None), (AL.inline_views(ldate), None)], 65536, rows_per_page)]
    tables = [O.OTable.from_pages(p, 65536, orc_cols(s)) for p, s in zip(pages, (CUSTOMER_SCHEMA, ORDERS_SCHEMA, LINEITEM_Q3_SCHEMA))]
    return pages, tables


def assert_q3_stream_equals(res, groups: dict, rel=1e-12):
    """res: pipeline result or oracle AggOut keyed (l_orderkey, o_orderdate, o_shippriority) -> (revenue,);
    groups: O.Q3Stream.groups()."""
    got = res.by_key()
    assert set(got) == set(groups), f"{len(got)} vs {len(groups)} groups"
    for k, (s, _) in groups.items():
        assert_close(got[k][0], s, rel, f"group {k}")
