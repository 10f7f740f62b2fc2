"""Result pages: pgf_result_encode_pages produces the page format that
ResultPageProducer emits (worker_runtime/src/result_pages.rs:150-196) and that ArrowPageDecoder
(page/import/src/lib.rs:117-206) accepts.  Checked with the oracle's restatement of import_owned
and with the product's own host-side import checks.  CPU only: the pgf_result is built by hand."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyorc as O
from pg_fusion_b200 import _lib
from pg_fusion_b200.worker import encode_result_pages

V_NULL, V_F64, V_I64, V_I128, V_STR = 0, 1, 2, 3, 4
T_I32, T_I64, T_F64, T_VIEW = 3, 4, 6, 8


def make_result(rows):
    """rows: list of (key_str|None, key_i32|None, sum_f64|None, count_i64)"""
    n = len(rows)
    keys = (_lib.Value * (2 * n + 1))()
    aggs = (_lib.Value * (2 * n + 1))()
    for g, (ks, ki, sf, cnt) in enumerate(rows):
        k0, k1, a0, a1 = keys[2 * g], keys[2 * g + 1], aggs[2 * g], aggs[2 * g + 1]
        if ks is None:
            k0.kind = V_NULL
        else:
            k0.kind, k0.slen = V_STR, len(ks)
            for i, b in enumerate(ks):
                k0.str[i] = b
        if ki is None:
            k1.kind = V_NULL
        else:
            k1.kind, k1.lo, k1.hi = V_I64, ki, -1 if ki < 0 else 0
        if sf is None:
            a0.kind = V_NULL
        else:
            a0.kind, a0.f64 = V_F64, sf
        a1.kind, a1.lo = V_I64, cnt
    r = _lib.Result()
    r.ngroups, r.nkeys, r.naggs = n, 2, 2
    r.keys, r.aggs = C.cast(keys, C.POINTER(_lib.Value)), C.cast(aggs, C.POINTER(_lib.Value))
    r.key_type[0], r.key_type[1] = T_VIEW, T_I32
    r.agg_type[0], r.agg_type[1] = T_F64, T_I64
    return r, (keys, aggs)


def sample_rows(n, seed=3):
    rng = np.random.default_rng(seed)
    rows = []
    for g in range(n):
        ks = None if g % 97 == 5 else bytes(rng.integers(65, 91, int(rng.integers(0, 13))).astype(np.uint8))
        ki = None if g % 89 == 7 else int(rng.integers(-2**31, 2**31))
        sf = None if g % 101 == 3 else float(rng.normal() * 1e6)
        rows.append((ks, ki, sf, int(rng.integers(0, 2**40))))
    return rows


@pytest.mark.parametrize("page_size,n", [(65536, 10_000), (4096, 1_000), (65536, 1), (8192, 0)])
def test_result_pages_decode_to_the_same_rows(page_size, n):
    rows = sample_rows(n)
    r, keep = make_result(rows)
    schema, pages = encode_result_pages(C.pointer(r), page_size)
    assert [(int(c.type_tag), c.nullable) for c in schema] == [(T_VIEW, True), (T_I32, True), (T_F64, True), (T_I64, False)]
    cols = [(int(c.type_tag), bool(c.nullable)) for c in schema]
    # the reference's fixed row cap decides the page count (page/row_estimator/src/lib.rs:353-371)
    cap = O.fixed_row_cap(cols, page_size - 20)
    assert pages.shape[0] == (n + cap - 1) // cap
    L = _lib.lib()
    specs = (_lib.ColumnSpec * 4)(*[_lib.ColumnSpec(int(c.type_tag), int(c.nullable)) for c in schema])
    for p in range(pages.shape[0]):
        page = np.ascontiguousarray(pages[p])
        kind, flags, plen = C.c_uint16(), C.c_uint16(), C.c_uint32()
        assert L.pgf_page_header_decode(page.ctypes.data_as(C.c_void_p), C.byref(kind), C.byref(flags), C.byref(plen)) == 0
        assert (kind.value, flags.value, plen.value) == (0x4152, 0, page_size - 20)
        block = page[20:]
        assert L.pgf_block_import_check(kind.value, flags.value, block.ctypes.data_as(C.c_void_p), block.size, specs, 4) == 0
        assert O.import_check(kind.value, flags.value, np.ascontiguousarray(block), cols) == 0
    if n == 0:
        return
    t = O.OTable.from_pages(pages, page_size, cols)   # oracle restatement of import_owned
    assert t.rows == n

    def values(i):
        c = t.column(i)
        if isinstance(c, list):
            return c
        arr, valid = c
        return [None if (valid is not None and not valid[r]) else arr[r] for r in range(len(arr))]
    c0, c1, c2, c3 = (values(i) for i in range(4))
    for g, (ks, ki, sf, cnt) in enumerate(rows):
        assert c0[g] == ks, (g, c0[g], ks)
        assert (None if c1[g] is None else int(c1[g])) == ki
        assert (None if c2[g] is None else float(c2[g])) == sf
        assert int(c3[g]) == cnt


def test_result_pages_with_integer_keys_and_decimal_sums():
    """The D variant's output: Int16 / Int64 group keys, Decimal128 SUM / AVG (16-byte slots, the
    extension tag 10), COUNT(*) -- every page passes the import checks and decodes to the same rows."""
    T_I16, T_DEC = 2, 10
    rng = np.random.default_rng(9)
    n = 3000
    keys = (_lib.Value * (2 * n + 1))()
    aggs = (_lib.Value * (2 * n + 1))()
    want = []
    for g in range(n):
        k16, k64 = int(rng.integers(-2**15, 2**15)), int(rng.integers(-2**62, 2**62))
        dec = None if g % 53 == 0 else int(rng.integers(-10**18, 10**18)) * int(rng.integers(1, 10**15))
        cnt = int(rng.integers(0, 2**50))
        keys[2 * g].kind, keys[2 * g].lo, keys[2 * g].hi = V_I64, k16, -1 if k16 < 0 else 0
        keys[2 * g + 1].kind, keys[2 * g + 1].lo, keys[2 * g + 1].hi = V_I64, k64, -1 if k64 < 0 else 0
        if dec is None:
            aggs[2 * g].kind = V_NULL
        else:
            u = dec & (2**128 - 1)
            lo, hi = u & (2**64 - 1), u >> 64
            aggs[2 * g].kind = V_I128
            aggs[2 * g].lo = lo - 2**64 if lo >= 2**63 else lo
            aggs[2 * g].hi = hi - 2**64 if hi >= 2**63 else hi
        aggs[2 * g + 1].kind, aggs[2 * g + 1].lo = V_I64, cnt
        want.append((k16, k64, dec, cnt))
    r = _lib.Result()
    r.ngroups, r.nkeys, r.naggs = n, 2, 2
    r.keys, r.aggs = C.cast(keys, C.POINTER(_lib.Value)), C.cast(aggs, C.POINTER(_lib.Value))
    r.key_type[0], r.key_type[1] = T_I16, T_I64
    r.agg_type[0], r.agg_type[1] = T_DEC, T_I64
    schema, pages = encode_result_pages(C.pointer(r), 65536)
    cols = [(int(c.type_tag), bool(c.nullable)) for c in schema]
    assert cols == [(T_I16, True), (T_I64, True), (T_DEC, True), (T_I64, False)]
    for p in range(pages.shape[0]):
        assert O.import_check(0x4152, 0, np.ascontiguousarray(pages[p, 20:]), cols) == 0
    t = O.OTable.from_pages(pages, 65536, cols)
    assert t.rows == n
    k0, v0 = t.column(0)
    k1, v1 = t.column(1)
    raw, vd = t.column(2)
    c3, _ = t.column(3)
    for g, (k16, k64, dec, cnt) in enumerate(want):
        assert int(k0[g]) == k16 and int(k1[g]) == k64 and int(c3[g]) == cnt
        if dec is None:
            assert vd is not None and not vd[g]
        else:
            assert (vd is None or vd[g]) and int.from_bytes(bytes(raw[g]), "little", signed=True) == dec


def test_nullability_follows_the_aggregate_function_not_the_data():
    """The receiving schema is fixed by the plan (page/import/src/lib.rs:150-180 rejects a nullability
    mismatch): SUM(Int64) is nullable even when this result holds no NULL, COUNT never is."""
    rows = [(b"a", 1, 1.5, 10), (b"b", 2, 2.5, 20)]
    r, keep = make_result(rows)
    r.agg_type[0], r.agg_type[1] = T_I64, T_I64
    for g in range(2):
        r.aggs[2 * g].kind, r.aggs[2 * g].lo = V_I64, 100 + g
    r.agg_func[0], r.agg_func[1] = 1, 3                      # SUM, COUNT(*)
    schema, pages = encode_result_pages(C.pointer(r), 65536)
    assert [(int(c.type_tag), c.nullable) for c in schema] == [(T_VIEW, True), (T_I32, True), (T_I64, True), (T_I64, False)]
    cols = [(int(c.type_tag), bool(c.nullable)) for c in schema]
    assert O.import_check(0x4152, 0, np.ascontiguousarray(pages[0][20:]), cols) == 0
    r.agg_func[0], r.agg_func[1] = 4, 2                      # COUNT(x), AVG
    schema, _ = encode_result_pages(C.pointer(r), 65536)
    assert [c.nullable for c in schema[2:]] == [False, True]


def test_encoding_can_continue_from_a_start_row():
    """page/batch_encoder/src/tests.rs:476-530 (append_batch_can_continue_from_start_row) through
    pgf_result_encode_pages(first_row, max_pages): pages produced one call at a time carry exactly the rows of
    a single call over the whole result, and a call that starts at the end produces nothing."""
    rows = sample_rows(5000, seed=9)
    r, keep = make_result(rows)
    schema, all_pages = encode_result_pages(C.pointer(r), 4096)
    cols = [(int(c.type_tag), bool(c.nullable)) for c in schema]
    cap = O.fixed_row_cap(cols, 4096 - 20)
    L = _lib.lib()
    first, stepped = 0, []
    while True:
        page = np.zeros(4096, dtype=np.uint8)
        got, done = C.c_uint64(), C.c_uint64()
        assert L.pgf_result_encode_pages(C.pointer(r), 4096, first, page.ctypes.data_as(C.c_void_p), 1, C.byref(got), C.byref(done)) == 0
        if got.value == 0:
            assert done.value == 0 and first == len(rows)
            break
        assert got.value == 1 and done.value == min(cap, len(rows) - first)     # full pages until the last one
        stepped.append(page)
        first += done.value
    assert len(stepped) == all_pages.shape[0] and all((a == b).all() for a, b in zip(stepped, all_pages))
    # starting in the middle of a page boundary-free position: rows [7, 7 + cap) land in one page
    page = np.zeros(4096, dtype=np.uint8)
    got, done = C.c_uint64(), C.c_uint64()
    assert L.pgf_result_encode_pages(C.pointer(r), 4096, 7, page.ctypes.data_as(C.c_void_p), 1, C.byref(got), C.byref(done)) == 0
    t = O.OTable.from_pages(page.reshape(1, 4096), 4096, cols)
    assert t.rows == cap and [int(c) for c in t.column(3)[0]] == [row[3] for row in rows[7:7 + cap]]
    # a start row beyond the result is an argument error
    assert L.pgf_result_encode_pages(C.pointer(r), 4096, len(rows) + 1, page.ctypes.data_as(C.c_void_p), 1, C.byref(got), C.byref(done)) == 1


def test_group_key_nullability_follows_the_source_column():
    """DataFusion's AggregateExec output field of a group expression keeps the nullability of its
    input field, and ArrowPageDecoder::validate_schema (page/import/src/lib.rs:208-235) rejects a page whose
    flags differ from the receiving schema (SchemaNullabilityMismatch).  key_not_null carries the plan's
    schema into the result; a page written with it decodes under the PLAN-derived schema and is refused under
    the other one."""
    rows = [(b"A", 1, 1.5, 10), (b"N", 2, 2.5, 20), (b"R", 3, None, 0)]
    r, keep = make_result(rows)
    r.agg_func[0], r.agg_func[1] = 1, 3                      # SUM, COUNT(*)
    r.key_not_null[0], r.key_not_null[1] = 1, 0              # l_returnflag NOT NULL, second key nullable
    schema, pages = encode_result_pages(C.pointer(r), 65536)
    plan_schema = [(T_VIEW, False), (T_I32, True), (T_F64, True), (T_I64, False)]   # what the planner derives
    assert [(int(c.type_tag), bool(c.nullable)) for c in schema] == plan_schema
    block = np.ascontiguousarray(pages[0][20:])
    assert O.import_check(0x4152, 0, block, plan_schema) == 0
    assert O.import_check(0x4152, 0, block, [(T_VIEW, True)] + plan_schema[1:]) != 0   # the old, all-nullable schema
    t = O.OTable.from_pages(pages, 65536, plan_schema)
    assert t.rows == 3 and t.column(0) == [b"A", b"N", b"R"]
    # a NULL key value in a column declared NOT NULL cannot be encoded
    r.keys[0].kind = V_NULL
    with pytest.raises(Exception):
        encode_result_pages(C.pointer(r), 65536)
