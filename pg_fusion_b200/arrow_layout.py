"""Host mirror of pg_fusion's `page/arrow_layout` + `page/transfer` page header + the
`page/import` checks, bound to libpgf_b200's host-side layout functions.

Same names and argument meaning as the reference (LayoutPlan::new plan.rs:33-93,
BlockRef::open access.rs:36-42, init_block access.rs:640-654, compute_fixed_row_cap
row_estimator/src/lib.rs:353-371, ArrowPageDecoder::import_owned import/src/lib.rs:117-206).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from enum import IntEnum
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .errors import PgfError

BLOCK_MAGIC = 0x32424150
PAGE_HEADER_LEN = 20
ARROW_LAYOUT_BATCH_KIND = 0x4152
DEFAULT_PAGE_SIZE = 65536                       # pg_fusion.page_size, pg/extension/src/guc.rs:31-32
DEFAULT_BLOCK_SIZE = DEFAULT_PAGE_SIZE - PAGE_HEADER_LEN  # 65516, row_encoder/benches/q05_encode.rs:8


class TypeTag(IntEnum):  # page/arrow_layout/src/types.rs:93-112
    Boolean = 1
    Int16 = 2
    Int32 = 3
    Int64 = 4
    Float32 = 5
    Float64 = 6
    Uuid = 7
    Utf8View = 8
    BinaryView = 9
    Decimal128 = 10  # extension (16-byte LE two's complement), not in reference v1 pages

    @property
    def is_view(self) -> bool:
        return self in (TypeTag.Utf8View, TypeTag.BinaryView)

    @property
    def row_width(self) -> int:
        return {2: 2, 3: 4, 5: 4, 4: 8, 6: 8, 7: 16, 8: 16, 9: 16, 10: 16}.get(int(self), 0)

    @property
    def numpy_dtype(self):
        return {2: np.int16, 3: np.int32, 4: np.int64, 5: np.float32, 6: np.float64}.get(int(self))


@dataclass(frozen=True)
class ColumnSpec:  # types.rs:218-225
    type_tag: TypeTag
    nullable: bool = False


def _specs(cols: Sequence[ColumnSpec]):
    arr = (_lib.ColumnSpec * max(1, len(cols)))()
    for i, c in enumerate(cols):
        arr[i].type_tag = int(c.type_tag)
        arr[i].nullable = 1 if c.nullable else 0
    return arr


def _check(rc: int, what: str = ""):
    if rc:
        raise PgfError(rc, what)


class LayoutPlan:
    """LayoutPlan::new (plan.rs:33-93)."""

    def __init__(self, specs: Sequence[ColumnSpec], max_rows: int, block_size: int):
        self.specs = list(specs)
        self.c = _lib.LayoutPlanC()
        _check(_lib.lib().pgf_layout_plan_new(_specs(specs), len(specs), max_rows, block_size, C.byref(self.c)),
               "LayoutPlan::new")

    @staticmethod
    def new(specs, max_rows, block_size) -> "LayoutPlan":
        return LayoutPlan(specs, max_rows, block_size)

    block_size = property(lambda self: self.c.block_size)
    max_rows = property(lambda self: self.c.max_rows)
    front_base = property(lambda self: self.c.front_base)
    pool_base = property(lambda self: self.c.pool_base)

    def column_layout(self, i: int):
        return self.c.cols[i]


def fixed_row_cap(specs: Sequence[ColumnSpec], block_size: int = DEFAULT_BLOCK_SIZE) -> int:
    cap = C.c_uint32()
    _check(_lib.lib().pgf_layout_fixed_row_cap(_specs(specs), len(specs), block_size, C.byref(cap)), "fixed_row_cap")
    return cap.value


LAYOUT_EXT_DECIMAL128 = 1


def validate_block(block: np.ndarray, extensions: int = 0) -> int:
    """BlockRef::open: returns the pgf_status (0 = valid).  Strict reference v1 (type tags 1..9) unless
    `extensions` switches this library's Decimal128 tag on."""
    block = np.ascontiguousarray(block, dtype=np.uint8)
    if extensions:
        return _lib.lib().pgf_block_validate_ext(block.ctypes.data_as(C.c_void_p), block.size, extensions)
    return _lib.lib().pgf_block_validate(block.ctypes.data_as(C.c_void_p), block.size)


def import_check(kind: int, flags: int, block: np.ndarray, schema: Sequence[ColumnSpec]) -> int:
    """All ArrowPageDecoder::import_owned checks on the host: returns the pgf_status."""
    block = np.ascontiguousarray(block, dtype=np.uint8)
    return _lib.lib().pgf_block_import_check(kind, flags, block.ctypes.data_as(C.c_void_p), block.size,
                                             _specs(schema), len(schema))


def page_header(kind: int = ARROW_LAYOUT_BATCH_KIND, flags: int = 0, payload_len: int = DEFAULT_BLOCK_SIZE) -> bytes:
    out = (C.c_uint8 * 20)()
    _check(_lib.lib().pgf_page_header_encode(kind, flags, payload_len, out))
    return bytes(out)


def inline_views(values: Sequence[Optional[bytes]]) -> np.ndarray:
    """ByteView::new_inline (raw.rs:114-126) for every value (<= 12 bytes): n x 16 bytes."""
    out = np.zeros((len(values), 16), dtype=np.uint8)
    for i, v in enumerate(values):
        if v is None:
            continue
        if len(v) > 12:
            raise ValueError("inline views hold at most 12 bytes")
        out[i, :4] = np.frombuffer(np.int32(len(v)).tobytes(), dtype=np.uint8)
        out[i, 4:4 + len(v)] = np.frombuffer(v, dtype=np.uint8)
    return out


def validity_bitmap(valid: Sequence[bool]) -> np.ndarray:
    """LSB-first validity bitmap (bitmap.rs:4-29)."""
    return np.packbits(np.asarray(valid, dtype=bool), bitorder="little")


def encode_block(specs: Sequence[ColumnSpec], columns: Sequence[Tuple[np.ndarray, Optional[np.ndarray]]],
                 nrows: int, max_rows: int, block_size: int = DEFAULT_BLOCK_SIZE) -> np.ndarray:
    """init_block + bulk column writes (the batch_encoder path): one arrow_layout block.

    columns[i] = (values, validity) where values is the raw value buffer (fixed width: numpy array of
    the column dtype; views/uuid/decimal: (nrows, 16) uint8) and validity an optional bool array."""
    L = _lib.lib()
    plan = LayoutPlan(specs, max_rows, block_size)
    block = np.zeros(block_size, dtype=np.uint8)
    bp = block.ctypes.data_as(C.c_void_p)
    _check(L.pgf_block_init(bp, block.size, C.byref(plan.c)), "init_block")
    for i, (vals, valid) in enumerate(columns):
        if int(specs[i].type_tag) == int(TypeTag.Boolean):   # bool array -> LSB-first bit-packed values (types.rs:139-147)
            vals = np.packbits(np.asarray(vals, dtype=bool), bitorder="little")
            if vals.size == 0:
                vals = np.zeros(1, np.uint8)
        vals = np.ascontiguousarray(vals)
        vb = None
        if valid is not None:
            vb = validity_bitmap(valid)
        _check(L.pgf_block_write_column(bp, block.size, i, nrows, vals.ctypes.data_as(C.c_void_p),
                                        None if vb is None else vb.ctypes.data_as(C.c_void_p)), "write_column")
    _check(L.pgf_block_set_row_count(bp, block.size, nrows), "set_row_count")
    return block


def encode_pages(specs: Sequence[ColumnSpec], columns: Sequence[Tuple[np.ndarray, Optional[np.ndarray]]],
                 page_size: int = DEFAULT_PAGE_SIZE, rows_per_page: Optional[int] = None) -> np.ndarray:
    """Split whole columns into transfer pages (20-byte header + block) of `page_size` bytes."""
    block_size = page_size - PAGE_HEADER_LEN
    cap = fixed_row_cap(specs, block_size)
    rpp = cap if rows_per_page is None else min(rows_per_page, cap)
    n = len(columns[0][0]) if columns else 0
    npages = max(1, -(-n // rpp)) if n else 0
    pages = np.zeros((npages, page_size), dtype=np.uint8)
    hdr = np.frombuffer(page_header(ARROW_LAYOUT_BATCH_KIND, 0, block_size), dtype=np.uint8)
    for p in range(npages):
        lo, hi = p * rpp, min(n, (p + 1) * rpp)
        cols = [(v[lo:hi], None if m is None else m[lo:hi]) for v, m in columns]
        pages[p, :20] = hdr
        pages[p, 20:] = encode_block(specs, cols, hi - lo, cap, block_size)
    return pages
