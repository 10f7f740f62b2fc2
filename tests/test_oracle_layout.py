"""Pins the page-layout oracle (oracle/orc_layout.c) against the reference's own tests:
page/arrow_layout/src/tests.rs and page/import/src/tests.rs (lines cited per test)."""
import ctypes as C
import struct

import numpy as np
import pytest

from oracle import pyorc as O

MIXED_COLS = [  # page/import/src/tests.rs:183-194 (all nullable)
    (O.T_BOOLEAN, True), (O.T_INT16, True), (O.T_INT32, True), (O.T_INT64, True), (O.T_FLOAT32, True),
    (O.T_FLOAT64, True), (O.T_UUID, True), (O.T_UTF8VIEW, True), (O.T_BINARYVIEW, True),
]
LONG_TXT = b"this string is definitely longer than twelve bytes"
LONG_BIN = b"this binary payload is also longer than twelve bytes"
MIXED_ROWS = [  # page/import/src/tests.rs:196-241
    (True, -7, 10, 100, 1.5, 3.5, bytes(range(1, 17)), b"short", b"\x01\x02"),
    (None,) * 9,
    (False, 9, 30, 300, -2.25, -4.75, bytes(range(16, 0, -1)), LONG_TXT, LONG_BIN),
    (True, 12, -40, -400, 0.0, 8.25, bytes(16), b"", b""),
]
FMT = {O.T_INT16: "<h", O.T_INT32: "<i", O.T_INT64: "<q", O.T_FLOAT32: "<f", O.T_FLOAT64: "<d"}


def encode_rows(cols, rows, block_size, max_rows=None):
    """Restates the reference's own minimal page writer `encode_layout_payload`
    (page/import/src/tests.rs:255-371): init_block -> write_* -> commit_current_row."""
    blk = O.Block(cols, len(rows) if max_rows is None else max_rows, block_size)
    for r, row in enumerate(rows):
        for c, (tag, _) in enumerate(cols):
            v = row[c]
            if v is None:
                blk.write_null(c, r)
            elif tag == O.T_BOOLEAN:
                blk.write_bool(c, r, v)
            elif tag in FMT:
                blk.write_fixed(c, r, struct.pack(FMT[tag], v))
            elif tag in (O.T_UUID, O.T_DECIMAL128):
                blk.write_fixed(c, r, v)
            else:
                assert blk.write_view_bytes(c, r, v) == 0
        blk.commit_current_row()
    assert blk.validate() == 0
    return blk


def test_repr_c_sizes_are_stable():
    # arrow_layout/src/tests.rs:7-17
    assert C.sizeof(O.ColumnLayout) == 20  # planner-side; the on-page structs are checked below
    hdr = struct.calcsize("<IHHIIIHHIIII")
    desc = struct.calcsize("<HHIIII")
    assert (hdr, desc) == (40, 20)


def test_plans_mixed_fixed_and_view_schema():
    # arrow_layout/src/tests.rs:19-45
    cols = [(O.T_BOOLEAN, True), (O.T_INT64, True), (O.T_UUID, False), (O.T_UTF8VIEW, True), (O.T_BINARYVIEW, True)]
    plan = O.layout_plan(cols, 64, 4096)
    assert plan.max_rows == 64 and plan.ncols == 5
    assert plan.front_base == 140
    assert plan.pool_base > plan.front_base
    assert plan.front_base % 16 == 12 and plan.pool_base % 16 == 12
    for i in range(5):
        l = plan.cols[i]
        assert l.validity_off % 16 == 12 and l.values_off % 16 == 12
        assert l.values_off >= l.validity_off + l.validity_len
    # exact offsets follow plan.rs:52-79: validity 16 B each; values 16, 512, 1024, 1024, 1024
    assert [plan.cols[i].values_len for i in range(5)] == [16, 512, 1024, 1024, 1024]
    assert plan.pool_base == 140 + 5 * 16 + 16 + 512 + 3 * 1024


def test_validates_header_and_column_descs_roundtrip():
    # arrow_layout/src/tests.rs:104-117
    blk = O.Block([(O.T_INT32, True), (O.T_UTF8VIEW, True)], 32, 2048)
    assert blk.validate() == 0
    hdr = struct.unpack_from("<IHHIIIHHIIII", blk.buf, 0)
    assert hdr[0] == 0x32424150 and hdr[1] == 1 and hdr[3] == 2048 and hdr[4] == 32 and hdr[5] == 0
    assert hdr[6] == 2 and hdr[8] == blk.plan.front_base and hdr[9] == blk.plan.pool_base and hdr[10] == 2048
    d0 = struct.unpack_from("<HHIIII", blk.buf, 40)
    d1 = struct.unpack_from("<HHIIII", blk.buf, 60)
    assert d0[:2] == (O.T_INT32, 1) and d1[:2] == (O.T_UTF8VIEW, 3)  # NULLABLE=1, VIEW=2 (types.rs:44-46)


def test_detects_inconsistent_view_flag():
    # arrow_layout/src/tests.rs:119-147
    blk = O.Block([(O.T_INT32, True), (O.T_UTF8VIEW, True)], 8, 1024)
    flags = struct.unpack_from("<H", blk.buf, 42)[0]
    struct.pack_into("<H", blk.buf, 42, flags | 2)
    assert blk.validate() == 110


def test_detects_too_small_block():
    # arrow_layout/src/tests.rs:149-166
    with pytest.raises(O.OracleError) as e:
        O.layout_plan([(O.T_INT64, True), (O.T_UTF8VIEW, True)], 128, 64)
    assert e.value.code == 113


def test_header_validation_errors():
    # page/arrow_layout/src/validate.rs:23-83
    def fresh():
        return O.Block([(O.T_INT32, False)], 8, 512)
    b = fresh(); struct.pack_into("<I", b.buf, 0, 0xDEADBEEF); assert b.validate() == 101
    b = fresh(); struct.pack_into("<H", b.buf, 4, 2); assert b.validate() == 102
    b = fresh(); struct.pack_into("<I", b.buf, 16, 9); assert b.validate() == 103  # row_count > max_rows
    b = fresh(); struct.pack_into("<I", b.buf, 24, 76); assert b.validate() == 105  # front_base
    b = fresh(); struct.pack_into("<I", b.buf, 32, 1024); assert b.validate() == 106  # tail_cursor > block_size
    b = fresh(); struct.pack_into("<I", b.buf, 28, b.plan.pool_base + 16); assert b.validate() == 112
    b = fresh(); struct.pack_into("<I", b.buf, 40 + 16, 1); assert b.validate() == 111  # reserved0 != 0
    assert O.block_validate(fresh().buf[:100]) == 108


def test_byte_view_inline_and_outline_round_trip():
    # arrow_layout/src/tests.rs:60-102
    blk = O.Block([(O.T_UTF8VIEW, True)], 4, 512)
    assert blk.write_view_bytes(0, 0, b"hello") == 0
    off = blk.plan.cols[0].values_off
    ln, data = struct.unpack_from("<i12s", blk.buf, off)
    assert ln == 5 and data == b"hello" + bytes(7)
    payload = b"abcdefghijklmnop"
    assert blk.write_view_bytes(0, 1, payload) == 0
    ln, prefix, idx, o = struct.unpack_from("<i4sii", blk.buf, off + 16)
    assert (ln, prefix, idx) == (16, b"abcd", 0)
    tail = struct.unpack_from("<I", blk.buf, 32)[0]
    assert tail == 512 - 16 and o == tail - blk.plan.pool_base
    assert bytes(blk.buf[tail:tail + 16]) == payload
    # ViewWriteStatus::Full when the tail arena is exhausted (access.rs:541-557)
    assert blk.write_view_bytes(0, 2, bytes(600)) in (122, 114)


def test_imports_mixed_batch():
    # page/import/src/tests.rs:373-393 -- the import checks pass and the decoded values equal the batch
    blk = encode_rows(MIXED_COLS, MIXED_ROWS, 4096)
    assert O.import_check(O.KIND_ARROW_LAYOUT, 0, blk.buf, MIXED_COLS) == 0
    page = np.zeros(4096 + 20, dtype=np.uint8)
    page[:20] = np.frombuffer(O.page_header(O.KIND_ARROW_LAYOUT, 0, 4096), dtype=np.uint8)
    page[20:] = blk.buf
    t = O.OTable.from_pages(page, page.size, MIXED_COLS)
    assert t.rows == 4
    for c, (tag, _) in enumerate(MIXED_COLS):
        want = [r[c] for r in MIXED_ROWS]
        if tag in (O.T_UTF8VIEW, O.T_BINARYVIEW):
            assert t.column(c) == want
            continue
        vals, valid = t.column(c)
        assert valid.tolist() == [1, 0, 1, 1]
        for r in (0, 2, 3):
            got = bytes(vals[r]) if tag == O.T_UUID else vals[r]
            assert got == want[r]
    # null counts recorded by commit_current_row (access.rs:443-457)
    for c in range(9):
        assert struct.unpack_from("<HHIIII", blk.buf, 40 + 20 * c)[4] == 1


def test_page_header_is_rmp_array_of_five():
    # page/transfer/src/page.rs:8-64
    h = O.page_header(0x4152, 0, 65516)
    assert len(h) == 20
    assert h == bytes([0x95, 0xCE, 0x50, 0x54, 0x50, 0x31, 0xCD, 0, 1, 0xCD, 0x41, 0x52, 0xCD, 0, 0, 0xCE, 0, 0, 0xFF, 0xEC])
    assert O.page_header_decode(h) == (0x4152, 0, 65516)
    with pytest.raises(O.OracleError):
        O.page_header_decode(b"\x94" + h[1:])


def test_import_rejections():
    blk = encode_rows(MIXED_COLS, MIXED_ROWS, 4096)
    # tests.rs:435-465
    assert O.import_check(9, 0, blk.buf, MIXED_COLS) == 201
    assert O.import_check(O.KIND_ARROW_LAYOUT, 1, blk.buf, MIXED_COLS) == 202
    # tests.rs:477-506
    bad = list(MIXED_COLS); bad[1] = (O.T_INT32, True)
    assert O.import_check(O.KIND_ARROW_LAYOUT, 0, blk.buf, bad) == 204
    assert O.import_check(O.KIND_ARROW_LAYOUT, 0, blk.buf, MIXED_COLS[:8]) == 203
    notnull = list(MIXED_COLS); notnull[2] = (O.T_INT32, False)
    assert O.import_check(O.KIND_ARROW_LAYOUT, 0, blk.buf, notnull) == 205
    # tests.rs:508-528: the reference test overwrites slot bytes 4..8 (the inline prefix of the long
    # value) and expects arrow's view validation to reject the page (ImportError::Arrow)
    b2 = encode_rows(MIXED_COLS, MIXED_ROWS, 4096)
    slot = b2.plan.cols[7].values_off + 2 * 16
    struct.pack_into("<i", b2.buf, slot + 4, 1)
    assert O.import_check(O.KIND_ARROW_LAYOUT, 0, b2.buf, MIXED_COLS) == 210
    # buffer_index != 0 in a long view (ByteView::validate, raw.rs:219-226)
    b2 = encode_rows(MIXED_COLS, MIXED_ROWS, 4096)
    struct.pack_into("<i", b2.buf, slot + 8, 1)
    assert O.import_check(O.KIND_ARROW_LAYOUT, 0, b2.buf, MIXED_COLS) == 118
    # non-zero padding after a short inline value / invalid UTF-8 in a Utf8View
    b2 = encode_rows(MIXED_COLS, MIXED_ROWS, 4096)
    b2.buf[b2.plan.cols[7].values_off + 4 + 9] = 1
    assert O.import_check(O.KIND_ARROW_LAYOUT, 0, b2.buf, MIXED_COLS) == 210
    b2 = encode_rows(MIXED_COLS, MIXED_ROWS, 4096)
    b2.buf[b2.plan.cols[7].values_off + 4] = 0xFF
    assert O.import_check(O.KIND_ARROW_LAYOUT, 0, b2.buf, MIXED_COLS) == 210
    b2.buf[b2.plan.cols[8].values_off + 4] = 0xFF  # BinaryView carries arbitrary bytes
    b2.buf[b2.plan.cols[7].values_off + 4] = ord("s")
    assert O.import_check(O.KIND_ARROW_LAYOUT, 0, b2.buf, MIXED_COLS) == 0
    # tests.rs:561-600: long view pointing before the allocated tail
    b3 = encode_rows(MIXED_COLS, MIXED_ROWS, 4096)
    struct.pack_into("<i", b3.buf, b3.plan.cols[7].values_off + 2 * 16 + 12, 0)
    assert O.import_check(O.KIND_ARROW_LAYOUT, 0, b3.buf, MIXED_COLS) == 208


def test_rejects_null_bitmap_count_mismatch():
    # tests.rs:530-559
    cols = [(O.T_BOOLEAN, True)]
    blk = encode_rows(cols, [(True,), (False,)], 512)
    blk.set_validity(0, 1, False)
    assert blk.validate() == 0
    assert O.import_check(O.KIND_ARROW_LAYOUT, 0, blk.buf, cols) == 207


def test_imports_empty_schema_batch():
    # tests.rs:420-433: zero columns, row_count 3
    blk = O.Block([], 3, 512)
    for _ in range(3):
        blk.commit_current_row()
    assert struct.unpack_from("<I", blk.buf, 16)[0] == 3
    assert O.import_check(O.KIND_ARROW_LAYOUT, 0, blk.buf, []) == 0


def test_slot_encoder_shape_payload():
    # tests.rs:602-711: (bool, i32, uuid) nullable, 3 rows with an all-NULL middle row
    cols = [(O.T_BOOLEAN, True), (O.T_INT32, True), (O.T_UUID, True)]
    rows = [(True, 11, bytes(range(1, 17))), (None, None, None), (False, -22, bytes(range(16, 0, -1)))]
    blk = encode_rows(cols, rows, 4096, max_rows=3)
    assert O.import_check(O.KIND_ARROW_LAYOUT, 0, blk.buf, cols) == 0


@pytest.mark.parametrize("cols,cap", [
    # SURVEY 8d rows/page table, block_size 65516 (page/row_encoder/benches/q05_encode.rs:8)
    ([(O.T_FLOAT64, False)] * 3 + [(O.T_UTF8VIEW, False)], 1614),
    ([(O.T_FLOAT64, False)] * 3 + [(O.T_INT32, False)], 2294),
    ([(O.T_FLOAT64, False)] * 4 + [(O.T_UTF8VIEW, False)] * 3, 806),
    ([(O.T_INT32, False), (O.T_FLOAT64, False), (O.T_FLOAT64, False), (O.T_UTF8VIEW, False)], 1791),
    ([(O.T_INT32, False), (O.T_INT32, False), (O.T_UTF8VIEW, False), (O.T_INT32, False)], 2293),
    ([(O.T_INT32, False), (O.T_UTF8VIEW, False)], 3229),
    ([(O.T_INT64, False)], 8056),
])
def test_fixed_row_cap_matches_survey_table(cols, cap):
    # page/row_estimator/src/lib.rs:353-371
    assert O.fixed_row_cap(cols, 65516) == cap
    O.layout_plan(cols, cap, 65516)
    with pytest.raises(O.OracleError):
        O.layout_plan(cols, cap + 1, 65516)


def test_row_estimator_fixed_width_cases():
    """page/row_estimator/src/tests.rs:21-65 for fixed-width shapes (the only ones the result encoder emits):
    the cap is exact (cap rows fit, cap + 1 do not), and a block that holds the header of a one-column layout
    but not one row has cap 0 (LayoutCannotFitAnyRows)."""
    cols = [(O.T_INT64, False), (O.T_BOOLEAN, True)]
    cap = O.fixed_row_cap(cols, 256)
    assert cap > 0
    O.layout_plan(cols, cap, 256)
    with pytest.raises(O.OracleError):
        O.layout_plan(cols, cap + 1, 256)
    one = [(O.T_INT64, False)]

    def fits(rows, size):
        try:
            O.layout_plan(one, rows, size)
            return True
        except O.OracleError:
            return False
    block_size = next(c for c in range(1, 512) if fits(0, c) and not fits(1, c))
    assert O.fixed_row_cap(one, block_size) == 0
