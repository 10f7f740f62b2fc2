"""oracle/pages_np.py -- TEST / BENCH INFRASTRUCTURE ONLY (see orc.h): numpy fabrication of TPC-H-shaped
tables as reference-format pages, using only the oracle's restatement of the layout rules
(oracle/orc_layout.c: LayoutPlan::new, init_block, page header).  `bench.py --impl reference` builds its
bounded samples with this, so that process never maps the product library.

Layout facts used (page/arrow_layout/src/raw.rs:21-46,69-84; plan.rs:33-93): block header 40 bytes with
row_count at byte 16; column descriptor c at 40 + 20 c with null_count at byte 12; values of column c at
values_off for max_rows x width bytes; NOT NULL columns only (null_count 0; the validity bit of every written
row is set, as the reference's writers do).
"""
from __future__ import annotations

import ctypes as C
import datetime as dt
from typing import List, Sequence, Tuple

import numpy as np

from . import pyorc as O

PAGE = 65536
HDR = 20
F64, I32, VIEW = O.T_FLOAT64, O.T_INT32, O.T_UTF8VIEW

Q6_COLS = [(F64, False)] * 3 + [(VIEW, False)]
Q1_COLS = [(F64, False)] * 4 + [(VIEW, False)] * 3
CUSTOMER_COLS = [(I32, False), (VIEW, False)]
ORDERS_COLS = [(I32, False), (I32, False), (VIEW, False), (I32, False)]
LINEITEM_Q3_COLS = [(I32, False), (F64, False), (F64, False), (VIEW, False)]

_WIDTH = {O.T_INT16: 2, O.T_INT32: 4, O.T_INT64: 8, O.T_FLOAT32: 4, O.T_FLOAT64: 8, O.T_UTF8VIEW: 16, O.T_BINARYVIEW: 16}


def inline_views(values: Sequence[bytes]) -> np.ndarray:
    """ByteView::new_inline (raw.rs:114-126): i32 length + up to 12 bytes, zero padded."""
    out = np.zeros((len(values), 16), dtype=np.uint8)
    lens = np.fromiter((len(v) for v in values), dtype=np.int32, count=len(values))
    assert lens.max(initial=0) <= 12
    out[:, :4] = lens.view(np.uint8).reshape(-1, 4)
    width = int(lens.max(initial=0))
    if width:
        padded = np.frombuffer(b"".join(v.ljust(width, b"\0") for v in values), dtype=np.uint8).reshape(-1, width)
        out[:, 4:4 + width] = padded
    return out


_DATES = None


def date_views(days: np.ndarray) -> np.ndarray:
    """ISO dates counted from 1992-01-01 as inline views (benches/tpch/schema.sql keeps dates as text)."""
    global _DATES
    if _DATES is None:
        base = dt.date(1992, 1, 1)
        _DATES = inline_views([(base + dt.timedelta(days=d)).isoformat().encode() for d in range(2600)])
    return _DATES[days]


def encode_pages(cols: Sequence[Tuple[int, bool]], columns: Sequence[np.ndarray], page_size: int = PAGE) -> np.ndarray:
    """columns[c]: n values (fixed width) or an (n, 16) uint8 array of views.  Returns [npages, page_size] uint8."""
    n = len(columns[0])
    block_size = page_size - HDR
    cap = O.fixed_row_cap(cols, block_size)
    plan = O.layout_plan(cols, cap, block_size)
    template = np.zeros(page_size, dtype=np.uint8)
    template[:HDR] = np.frombuffer(O.page_header(O.KIND_ARROW_LAYOUT, 0, block_size), dtype=np.uint8)
    block = template[HDR:]
    rc = O.lib().orc_init_block(block.ctypes.data_as(C.c_void_p), block.size, C.byref(plan))
    if rc:
        raise O.OracleError(rc, "init_block")
    npages = max(1, (n + cap - 1) // cap)
    pages = np.tile(template, (npages, 1))
    full, rest = divmod(n, cap)
    for c, (tag, _) in enumerate(cols):
        w = _WIDTH[tag]
        off = HDR + int(plan.cols[c].values_off)
        raw = np.ascontiguousarray(columns[c]).view(np.uint8).reshape(n, w)
        if full:
            pages[:full, off:off + cap * w] = raw[:full * cap].reshape(full, cap * w)
        if rest:
            pages[full, off:off + rest * w] = raw[full * cap:].reshape(rest * w)
    counts = np.full(npages, cap, dtype=np.uint32)
    if rest or n == 0:
        counts[-1] = rest
    # the reference's writers set the validity bit of every written row, nullable column or not (access.rs:320-323)
    for c in range(len(cols)):
        voff = HDR + int(plan.cols[c].validity_off)
        for p, rows in ((slice(0, full), cap), (slice(full, full + 1), rest)):
            if rows == 0 or (p.stop <= p.start):
                continue
            pages[p, voff:voff + rows // 8] = 0xFF
            if rows % 8:
                pages[p, voff + rows // 8] = (1 << (rows % 8)) - 1
    pages[:, HDR + 16:HDR + 20] = counts.view(np.uint8).reshape(npages, 4)
    for p in (0, npages - 1):   # the writer is checked by the oracle's own validators
        blk = np.ascontiguousarray(pages[p, HDR:])
        assert O.block_validate(blk) == 0 and O.import_check(O.KIND_ARROW_LAYOUT, 0, blk, cols) == 0
    return pages


def lineitem_columns(n: int, seed: int):
    """TPC-H-shaped lineitem: money as f64 = cents / 100.0, dates uniform over 2526 days, flags correlated."""
    r = np.random.default_rng(seed)
    qty = r.integers(1, 51, n)
    part = r.integers(90000, 210001, n)
    ship = r.integers(1, 2527, n)
    receipt = ship + r.integers(1, 31, n)
    cutoff = 1263  # 1995-06-17
    ls = np.where(ship > cutoff, ord("O"), ord("F")).astype(np.uint8)
    rf = np.where(receipt <= cutoff, np.where(r.integers(0, 2, n) == 1, ord("R"), ord("A")), ord("N")).astype(np.uint8)
    return dict(qty=qty.astype(np.float64), price=(qty * part) / 100.0, disc=r.integers(0, 11, n) / 100.0,
                tax=r.integers(0, 9, n) / 100.0, ship=ship, rf=rf, ls=ls)


def _char_views(codes: np.ndarray) -> np.ndarray:
    out = np.zeros((codes.size, 16), dtype=np.uint8)
    out[:, 0] = 1
    out[:, 4] = codes
    return out


def q6_pages(li) -> np.ndarray:
    return encode_pages(Q6_COLS, [li["qty"], li["price"], li["disc"], date_views(li["ship"])])


def q1_pages(li) -> np.ndarray:
    return encode_pages(Q1_COLS, [li["qty"], li["price"], li["disc"], li["tax"], _char_views(li["rf"]), _char_views(li["ls"]),
                                  date_views(li["ship"])])


def q3_pages(nli: int, seed: int):
    """customer / orders / lineitem in TPC-H proportions (1 : 10 : 40) with the correlations that decide Q3's
    selectivities, as in the device generator (pg_fusion_b200/csrc/gen.cu): a fifth of the customers is BUILDING,
    o_custkey uniform over the customers with custkey % 3 != 0, sparse order keys (8 of every 32 values), every
    lineitem picks an order uniformly and ships 1..121 days after the order date."""
    r = np.random.default_rng(seed)
    nord = max(10, nli // 4)
    ncust = max(10, nord // 10)
    segs = inline_views([b"AUTOMOBILE", b"BUILDING", b"FURNITURE", b"MACHINERY", b"HOUSEHOLD"])
    cust = encode_pages(CUSTOMER_COLS, [np.arange(1, ncust + 1, dtype=np.int32), segs[r.integers(0, 5, ncust)]])
    oi = np.arange(nord, dtype=np.int64)
    okey = ((oi // 8) * 32 + oi % 8 + 1).astype(np.int32)
    ck = r.integers(1, ncust + 1, nord)
    ck = np.where(ck % 3 == 0, np.where(ck > 1, ck - 1, ck + 1), ck).astype(np.int32)
    odays = r.integers(0, 2406, nord)
    orders = encode_pages(ORDERS_COLS, [okey, ck, date_views(odays), np.zeros(nord, dtype=np.int32)])
    pick = r.integers(0, nord, nli)
    li = encode_pages(LINEITEM_Q3_COLS, [okey[pick], r.integers(90000, 10_000_000, nli) / 100.0, r.integers(0, 11, nli) / 100.0,
                                         date_views(odays[pick] + r.integers(1, 122, nli))])
    return cust, orders, li
