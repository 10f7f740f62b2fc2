"""CPU tests of the product's host-side layout code (C++ in libpgf_b200.so) against the
oracle, and of the C-ABI surface itself.  No GPU needed: no compute entry point is called."""
import ctypes as C
import os
import re
import struct

import numpy as np
import pytest

import pg_fusion_b200 as pg
from oracle import pyorc as O
from pg_fusion_b200 import ColumnSpec, TypeTag, _lib
from pg_fusion_b200 import arrow_layout as AL

from . import util as U

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "pgf_b200.h")).read()
    declared = set(re.findall(r"\b(pgf_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"pgf_status"}
    L = _lib.lib()
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, f"declared in include/pgf_b200.h but not exported: {missing}"
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)


def test_no_device_means_no_context_and_no_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert pg.device_count() == 0
    with pytest.raises(pg.PgfError) as e:
        pg.Context()
    assert e.value.name == "NO_DEVICE"


SCHEMAS = [
    [ColumnSpec(TypeTag.Boolean, True), ColumnSpec(TypeTag.Int64, True), ColumnSpec(TypeTag.Uuid, False),
     ColumnSpec(TypeTag.Utf8View, True), ColumnSpec(TypeTag.BinaryView, True)],
    U.Q6_SCHEMA, U.Q1_SCHEMA,
    [ColumnSpec(TypeTag.Int32), ColumnSpec(TypeTag.Int32), ColumnSpec(TypeTag.Utf8View), ColumnSpec(TypeTag.Int32)],
    [ColumnSpec(TypeTag.Int16, True), ColumnSpec(TypeTag.Float32), ColumnSpec(TypeTag.Decimal128, True)],
    [],
]


@pytest.mark.parametrize("schema", SCHEMAS)
@pytest.mark.parametrize("max_rows,block_size", [(0, 4096), (64, 4096), (1, 65516), (700, 65516)])
def test_layout_plan_matches_oracle(schema, max_rows, block_size):
    try:
        want = O.layout_plan(U.orc_cols(schema), max_rows, block_size)
    except O.OracleError as e:
        with pytest.raises(pg.PgfError) as pe:
            AL.LayoutPlan(schema, max_rows, block_size)
        assert pe.value.code == e.code
        return
    got = AL.LayoutPlan(schema, max_rows, block_size)
    assert (got.front_base, got.pool_base, got.max_rows, got.block_size) == (want.front_base, want.pool_base, want.max_rows, want.block_size)
    for i in range(len(schema)):
        g, w = got.column_layout(i), want.cols[i]
        assert (g.type_tag, g.flags, g.validity_off, g.values_off, g.validity_len, g.values_len) == \
               (w.type_tag, w.flags, w.validity_off, w.values_off, w.validity_len, w.values_len)


@pytest.mark.parametrize("schema", [s for s in SCHEMAS if s])
def test_fixed_row_cap_matches_oracle(schema):
    assert AL.fixed_row_cap(schema, 65516) == O.fixed_row_cap(U.orc_cols(schema), 65516)


def test_page_header_matches_oracle():
    for kind, flags, n in [(0x4152, 0, 65516), (9, 1, 0), (0xFFFF, 0xFFFF, 2**32 - 1)]:
        assert AL.page_header(kind, flags, n) == O.page_header(kind, flags, n)


def _mixed_block_via_product():
    schema = [ColumnSpec(TypeTag.Int16, True), ColumnSpec(TypeTag.Int32, True), ColumnSpec(TypeTag.Int64, True),
              ColumnSpec(TypeTag.Float32, True), ColumnSpec(TypeTag.Float64, True), ColumnSpec(TypeTag.Utf8View, True)]
    valid = np.array([True, False, True, True])
    cols = [(np.array([-7, 0, 9, 12], np.int16), valid), (np.array([10, 0, 30, -40], np.int32), valid),
            (np.array([100, 0, 300, -400], np.int64), valid), (np.array([1.5, 0, -2.25, 0.0], np.float32), valid),
            (np.array([3.5, 0, -4.75, 8.25], np.float64), valid),
            (AL.inline_views([b"short", None, b"twelve bytes", b""]), valid)]
    return schema, cols, AL.encode_block(schema, cols, 4, 4, 4096)


def test_product_writer_is_byte_identical_to_oracle_writer():
    schema, cols, block = _mixed_block_via_product()
    ob = O.Block(U.orc_cols(schema), 4, 4096)
    fmt = {TypeTag.Int16: "<h", TypeTag.Int32: "<i", TypeTag.Int64: "<q", TypeTag.Float32: "<f", TypeTag.Float64: "<d"}
    strs = [b"short", None, b"twelve bytes", b""]
    for r in range(4):
        for c, spec in enumerate(schema):
            if r == 1:
                ob.write_null(c, r)
            elif spec.type_tag == TypeTag.Utf8View:
                assert ob.write_view_bytes(c, r, strs[r]) == 0
            else:
                ob.write_fixed(c, r, struct.pack(fmt[spec.type_tag], cols[c][0][r]))
        ob.commit_current_row()
    assert bytes(block) == bytes(ob.buf)
    assert AL.validate_block(block) == 0 and O.block_validate(block) == 0
    assert AL.import_check(0x4152, 0, block, schema) == 0
    assert O.import_check(0x4152, 0, block, U.orc_cols(schema)) == 0


def test_decimal128_extension_tag_is_rejected_by_the_reference_v1_checks():
    """TypeTag::from_raw knows 1..9 (page/arrow_layout/src/types.rs:93-112): BlockRef::open on a block carrying this
    library's Decimal128 tag is InvalidTypeTag, and so is an import under a schema that does not name the tag."""
    schema = [ColumnSpec(TypeTag.Int32), ColumnSpec(TypeTag.Decimal128, True)]
    vals = np.zeros((3, 16), np.uint8)
    vals[:, 0] = [1, 2, 3]
    block = AL.encode_block(schema, [(np.array([1, 2, 3], np.int32), None), (vals, np.array([True, False, True]))], 3, 3, 4096)
    invalid_tag = 109   # PGF_ERR_LAYOUT_INVALID_TYPE_TAG = LayoutError::InvalidTypeTag
    assert AL.validate_block(block) == O.block_validate_v1(block) == invalid_tag
    assert AL.validate_block(block, AL.LAYOUT_EXT_DECIMAL128) == O.block_validate(block) == 0
    # the caller's schema is the opt-in: it passes under the schema that names the tag ...
    assert AL.import_check(0x4152, 0, block, schema) == O.import_check(0x4152, 0, block, U.orc_cols(schema)) == 0
    # ... and is InvalidTypeTag (before any schema comparison) under a reference-only schema
    ref_schema = [ColumnSpec(TypeTag.Int32), ColumnSpec(TypeTag.Uuid, True)]
    assert AL.import_check(0x4152, 0, block, ref_schema) == O.import_check(0x4152, 0, block, U.orc_cols(ref_schema)) == invalid_tag


def test_import_rejections_match_oracle_codes():
    schema, cols, block = _mixed_block_via_product()
    oc = U.orc_cols(schema)

    def both(buf, kind=0x4152, flags=0, sch=schema):
        a = AL.import_check(kind, flags, buf, sch)
        b = O.import_check(kind, flags, np.ascontiguousarray(buf), U.orc_cols(sch))
        assert a == b, (a, b)
        return a

    assert both(block, kind=9) == 201
    assert both(block, flags=1) == 202
    assert both(block, sch=schema[:5]) == 203
    bad = list(schema); bad[1] = ColumnSpec(TypeTag.Int64, True)
    assert both(block, sch=bad) == 204
    bad = list(schema); bad[1] = ColumnSpec(TypeTag.Int32, False)
    assert both(block, sch=bad) == 205
    for off, val, code in [(0, 0xDEADBEEF, 101), (16, 99, 103), (24, 4, 105), (32, 99999, 106)]:
        b = block.copy(); struct.pack_into("<I", b, off, val); assert both(b) == code
    b = block.copy(); struct.pack_into("<H", b, 4, 7); assert both(b) == 102
    b = block.copy(); struct.pack_into("<I", b, 40 + 12, 3); assert both(b) == 207   # null_count lies
    b = block.copy(); struct.pack_into("<I", b, 40 + 12, 9); assert both(b) == 206   # null_count > rows
    plan = AL.LayoutPlan(schema, 4, 4096)
    voff = plan.column_layout(5).values_off
    b = block.copy(); b[voff + 4 + 9] = 1; assert both(b) == 210                      # inline padding
    b = block.copy(); b[voff + 4] = 0xFF; assert both(b) == 210                       # invalid UTF-8
    b = block.copy(); struct.pack_into("<i", b, voff, -1); assert both(b) == 117      # negative length
    b = block.copy(); struct.pack_into("<iiii", b, voff, 50, 0, 0, 5000); assert both(b) == 120
    assert both(block[:100]) == 108


def test_encode_pages_roundtrip_through_oracle_decoder():
    li = U.lineitem(5000, 7)
    pages = U.q1_pages(li)
    assert pages.shape == (7, 65536)  # 806 rows per page
    t = O.OTable.from_pages(pages, 65536, U.orc_cols(U.Q1_SCHEMA))
    assert t.rows == 5000
    got, _ = t.column(1)
    assert (got == li["price"]).all()
    assert t.column(6) == U.dates_from_days(li["ship"])
    assert t.column(4) == [bytes(x) for x in li["rf"]]


def test_validators_agree_with_the_oracle_on_mutated_blocks_under_sanitizers(tmp_path):
    """Differential fuzzing (tests/cpp/fuzz_layout.cpp): the product's page validators and the oracle's
    restatement of BlockRef::open / import_owned see 60 000 mutated blocks in exactly-sized heap buffers,
    built with -fsanitize=address,undefined.  Same status code for every block, no out-of-bounds read."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    san = ["-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer"]
    objs = []
    for cc, std, src in (("gcc", "-std=c11", "oracle/orc_layout.c"), ("g++", "-std=c++17", "pg_fusion_b200/csrc/layout.cpp")):
        obj = str(tmp_path / (os.path.basename(src) + ".o"))
        out = subprocess.run([cc, std, *san, "-I", os.path.join(root, "include"), "-c", os.path.join(root, src), "-o", obj],
                             capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        objs.append(obj)
    exe = str(tmp_path / "fuzz_layout")
    out = subprocess.run(["g++", "-std=c++17", *san, "-Wall", "-Wextra", "-I", os.path.join(root, "include"),
                          os.path.join(root, "tests", "cpp", "fuzz_layout.cpp"), *objs, "-o", exe], capture_output=True, text=True)
    if out.returncode != 0 and "asan" in out.stderr.lower():
        pytest.skip("sanitizer runtime not available")
    assert out.returncode == 0, out.stderr
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1", UBSAN_OPTIONS="halt_on_error=1")
    for seed in (1, 2):
        run = subprocess.run([exe, "30000", str(seed)], capture_output=True, text=True, timeout=600, env=env)
        assert run.returncode == 0 and "runtime error" not in run.stderr and "ERROR" not in run.stderr, run.stdout[-2000:] + run.stderr[-3000:]
        codes = {int(kv.split(":")[0]) for kv in [ln for ln in run.stdout.splitlines() if ln.startswith("codes")][0].split()[1:]}
        # every rejection reason of the validator and the import list has been exercised
        assert {101, 102, 103, 105, 106, 107, 108, 109, 110, 111, 112, 117, 118, 119, 120,
                201, 202, 203, 204, 205, 206, 207, 208, 210} <= codes, sorted(codes)


def test_row_estimator_fixed_width_cases_in_the_product():
    """The same cases as tests/test_oracle_layout.py::test_row_estimator_fixed_width_cases through the C ABI, and
    what the result encoder makes of them: no column -> no transport schema, no row fits -> DoesNotFit."""
    import ctypes as C
    from pg_fusion_b200 import _lib
    schema = [ColumnSpec(TypeTag.Int64), ColumnSpec(TypeTag.Boolean, True)]
    cap = AL.fixed_row_cap(schema, 256)
    assert cap == O.fixed_row_cap(U.orc_cols(schema), 256) > 0
    AL.LayoutPlan(schema, cap, 256)
    with pytest.raises(pg.PgfError):
        AL.LayoutPlan(schema, cap + 1, 256)
    one = [ColumnSpec(TypeTag.Int64)]
    sizes = [c for c in range(1, 512) if _fits(one, 0, c) and O.fixed_row_cap(U.orc_cols(one), c) == 0]
    assert sizes and all(AL.fixed_row_cap(one, c) == 0 for c in sizes)
    L = _lib.lib()
    r = _lib.Result()                                   # an aggregate result without any column
    n = C.c_uint32()
    assert L.pgf_result_schema(C.byref(r), (_lib.ColumnSpec * 20)(), C.byref(n)) == 1   # PGF_ERR_INVALID_ARGUMENT
    keys = (_lib.Value * 2)()
    keys[0].kind, keys[0].lo = 2, 7
    r.ngroups, r.nkeys, r.keys, r.aggs = 1, 1, C.cast(keys, C.POINTER(_lib.Value)), C.cast(keys, C.POINTER(_lib.Value))
    r.key_type[0] = int(TypeTag.Int64)
    got, rows = C.c_uint64(), C.c_uint64()
    buf = (C.c_uint8 * 4096)()
    page_size = sizes[0] + 20                           # holds the block header of the layout, but not one row
    assert L.pgf_result_encode_pages(C.byref(r), page_size, 0, buf, 1, C.byref(got), C.byref(rows)) in (1, 113)


def _fits(schema, rows, size):
    try:
        AL.LayoutPlan(schema, rows, size)
        return True
    except pg.PgfError:
        return False
