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
