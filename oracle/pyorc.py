"""ctypes binding for the CPU oracle (oracle/liborc.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (pg_fusion_b200) never imports
this module.  See oracle/orc.h for the reference file:line each function restates.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_DIR, "liborc.so")


def build(force: bool = False) -> str:
    """Compile oracle/liborc.so with gcc (plain C, no GPU)."""
    srcs = [os.path.join(_DIR, f) for f in ("orc_bloom.c", "orc_layout.c", "orc_ops.c", "orc_fast.c", "orc_q3.c", "orc.h")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    if stale:
        subprocess.check_call(["make", "-C", _DIR, "-B", "liborc.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


# type tags (page/arrow_layout/src/types.rs:93-112) + Decimal128 extension
T_BOOLEAN, T_INT16, T_INT32, T_INT64, T_FLOAT32, T_FLOAT64, T_UUID, T_UTF8VIEW, T_BINARYVIEW, T_DECIMAL128 = range(1, 11)
KIND_ARROW_LAYOUT = 0x4152
PAGE_HEADER_LEN = 20

# expression opcodes / aggregate funcs / value kinds (orc.h)
X_COL, X_LIT_F64, X_LIT_I64, X_LIT_STR, X_LIT_I128 = 1, 2, 3, 4, 5
X_ADD, X_SUB, X_MUL = 10, 11, 12
X_LT, X_LE, X_GT, X_GE, X_EQ, X_NE = 20, 21, 22, 23, 24, 25
X_AND = 30
AGG_SUM, AGG_AVG, AGG_COUNT_STAR, AGG_COUNT, AGG_MIN, AGG_MAX = 1, 2, 3, 4, 5, 6
V_NULL, V_F64, V_I64, V_I128, V_STR, V_BOOL = 0, 1, 2, 3, 4, 5

RF_FREE, RF_BUILDING, RF_READY, RF_DISABLED = 0, 1, 2, 3
PASS_UNFILTERED, MAYBE_PRESENT, DEFINITELY_ABSENT = 0, 1, 2


class BloomParams(C.Structure):
    _fields_ = [("bit_count", C.c_uint64), ("word_count", C.c_uint64), ("hash_count", C.c_uint64), ("seed", C.c_uint64)]


class ColumnSpec(C.Structure):
    _fields_ = [("type_tag", C.c_uint16), ("nullable", C.c_uint16)]


class ColumnLayout(C.Structure):
    _fields_ = [("type_tag", C.c_uint16), ("flags", C.c_uint16), ("validity_off", C.c_uint32),
                ("values_off", C.c_uint32), ("validity_len", C.c_uint32), ("values_len", C.c_uint32)]


class LayoutPlan(C.Structure):
    _fields_ = [("block_size", C.c_uint32), ("max_rows", C.c_uint32), ("front_base", C.c_uint32),
                ("pool_base", C.c_uint32), ("ncols", C.c_uint32), ("cols", ColumnLayout * 64)]


class Column(C.Structure):
    _fields_ = [("type_tag", C.c_int32), ("nullable", C.c_int32), ("rows", C.c_uint64),
                ("values", C.c_void_p), ("validity", C.c_void_p), ("arena", C.c_void_p)]


class Table(C.Structure):
    _fields_ = [("ncols", C.c_uint32), ("rows", C.c_uint64), ("cols", Column * 64)]


class XNode(C.Structure):
    _fields_ = [("op", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("slen", C.c_int32),
                ("f", C.c_double), ("i", C.c_int64), ("i2", C.c_int64), ("s", C.c_char * 16)]


class AggSpec(C.Structure):
    _fields_ = [("func", C.c_int32), ("expr_off", C.c_int32), ("expr_len", C.c_int32)]


class Value(C.Structure):
    _fields_ = [("kind", C.c_int32), ("slen", C.c_int32), ("f", C.c_double), ("lo", C.c_int64),
                ("hi", C.c_int64), ("s", C.c_char * 16)]


class JoinEdge(C.Structure):
    _fields_ = [("build", C.POINTER(Table)), ("build_col", C.c_int32), ("probe_src", C.c_int32),
                ("probe_col", C.c_int32)]


class AggResult(C.Structure):
    _fields_ = [("ngroups", C.c_uint64), ("nkeys", C.c_uint32), ("naggs", C.c_uint32),
                ("keys", C.POINTER(Value)), ("aggs", C.POINTER(Value)), ("rows_in", C.c_uint64),
                ("rows_filtered", C.c_uint64), ("rows_joined", C.c_uint64)]


class Q6Result(C.Structure):
    _fields_ = [("sum", C.c_double), ("rows_in", C.c_uint64), ("rows_kept", C.c_uint64)]


class Q1Group(C.Structure):
    _fields_ = [("returnflag", C.c_char), ("linestatus", C.c_char), ("sum_qty", C.c_double),
                ("sum_base_price", C.c_double), ("sum_disc_price", C.c_double), ("sum_charge", C.c_double),
                ("sum_disc", C.c_double), ("count", C.c_uint64)]


class Q1Result(C.Structure):
    _fields_ = [("ngroups", C.c_uint32), ("groups", Q1Group * 16), ("rows_in", C.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        u64, i32, u32, vp = C.c_uint64, C.c_int32, C.c_uint32, C.c_void_p
        P = C.POINTER
        L.orc_splitmix64.restype = u64
        L.orc_splitmix64.argtypes = [u64]
        L.orc_hash_int_key.restype = u64
        L.orc_hash_int_key.argtypes = [C.c_int64]
        L.orc_bloom_params_new.argtypes = [u64, u64, u64, P(BloomParams)]
        L.orc_bloom_params_for_expected_items.argtypes = [u64, C.c_double, u64, P(BloomParams)]
        L.orc_bloom_attach_check.argtypes = [P(BloomParams), vp, u64]
        L.orc_bloom_bit_index.restype = u64
        L.orc_bloom_bit_index.argtypes = [P(BloomParams), u64, u64]
        L.orc_bloom_clear.restype = None
        L.orc_bloom_clear.argtypes = [P(BloomParams), vp]
        L.orc_bloom_insert_hash.restype = None
        L.orc_bloom_insert_hash.argtypes = [P(BloomParams), vp, u64]
        L.orc_bloom_might_contain_hash.argtypes = [P(BloomParams), vp, u64]
        L.orc_bloom_insert_keys.restype = u64
        L.orc_bloom_insert_keys.argtypes = [P(BloomParams), vp, vp, C.c_int, vp, u64]
        L.orc_bloom_probe_keys.restype = u64
        L.orc_bloom_probe_keys.argtypes = [P(BloomParams), vp, vp, C.c_int, vp, u64, vp]
        L.orc_lifecycle_pack.argtypes = [u64, C.c_int, P(u64)]
        L.orc_lifecycle_unpack.restype = None
        L.orc_lifecycle_unpack.argtypes = [u64, P(u64), P(C.c_int)]
        L.orc_slot_try_acquire_builder.argtypes = [P(u64), P(BloomParams), vp, P(u64)]
        L.orc_slot_publish_build.argtypes = [P(u64), u64]
        L.orc_slot_disable_build.argtypes = [P(u64), u64]
        L.orc_slot_retire_ready.argtypes = [P(u64), u64]
        L.orc_probe_decision_for_hash.argtypes = [P(u64), u64, P(BloomParams), vp, u64]
        L.orc_probe_decision_for_null.argtypes = [P(u64), u64]
        L.orc_type_row_width.argtypes = [C.c_int]
        L.orc_layout_plan_new.argtypes = [P(ColumnSpec), u32, u32, u32, P(LayoutPlan)]
        L.orc_fixed_row_cap.argtypes = [P(ColumnSpec), u32, u32, P(u32)]
        L.orc_init_block.argtypes = [vp, C.c_size_t, P(LayoutPlan)]
        L.orc_block_validate.argtypes = [vp, C.c_size_t]
        L.orc_block_validate_v1.argtypes = [vp, C.c_size_t]
        L.orc_block_write_fixed.argtypes = [vp, C.c_size_t, u32, u32, vp, u32]
        L.orc_block_write_bool.argtypes = [vp, C.c_size_t, u32, u32, C.c_int]
        L.orc_block_write_null.argtypes = [vp, C.c_size_t, u32, u32]
        L.orc_block_write_view_bytes.argtypes = [vp, C.c_size_t, u32, u32, vp, u32]
        L.orc_block_commit_current_row.argtypes = [vp, C.c_size_t]
        L.orc_block_set_validity.argtypes = [vp, C.c_size_t, u32, u32, C.c_int]
        L.orc_page_header_encode.argtypes = [C.c_uint16, C.c_uint16, u32, vp]
        L.orc_page_header_decode.argtypes = [vp, P(C.c_uint16), P(C.c_uint16), P(u32)]
        L.orc_import_check.argtypes = [C.c_uint16, C.c_uint16, vp, C.c_size_t, P(ColumnSpec), u32]
        L.orc_table_from_pages.argtypes = [vp, u64, u64, P(ColumnSpec), u32, P(Table)]
        L.orc_table_free.restype = None
        L.orc_table_free.argtypes = [P(Table)]
        L.orc_table_select.argtypes = [P(Table), vp, P(Table)]
        L.orc_table_take.argtypes = [P(Table), vp, u64, P(Table)]
        L.orc_aggregate.argtypes = [P(Table), P(XNode), i32, i32, P(JoinEdge), u32, P(i32), P(i32), u32,
                                    P(AggSpec), u32, i32, i32, P(AggResult)]
        L.orc_agg_result_free.restype = None
        L.orc_agg_result_free.argtypes = [P(AggResult)]
        L.orc_filter.argtypes = [P(Table), P(XNode), i32, i32, vp, P(u64)]
        L.orc_hash_join_pairs.argtypes = [P(Table), i32, P(Table), i32, vp, vp, u64, P(u64)]
        L.orc_q6_pages.argtypes = [vp, u64, u64, C.c_int, P(i32), C.c_char_p, C.c_char_p, C.c_double,
                                   C.c_double, C.c_double, P(Q6Result)]
        L.orc_q1_pages.argtypes = [vp, u64, u64, C.c_int, P(i32), C.c_char_p, C.c_int, P(Q1Result)]
        L.orc_q3_new.restype = vp
        L.orc_q3_new.argtypes = [C.c_char_p, C.c_char_p]
        L.orc_q3_free.restype = None
        L.orc_q3_free.argtypes = [vp]
        for fn in (L.orc_q3_customer, L.orc_q3_orders, L.orc_q3_lineitem):
            fn.argtypes = [vp, vp, u64, u64, C.c_int]
        L.orc_q3_customer_finish.argtypes = [vp]
        L.orc_q3_orders_finish.argtypes = [vp, C.c_int]
        L.orc_q3_stats.argtypes = [vp, P(u64)]
        L.orc_q3_groups.restype = u64
        L.orc_q3_groups.argtypes = [vp, vp, vp, vp, vp, vp, vp, u64]
        _lib = L
    return _lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# --------------------------------------------------------------------------- Bloom

class OracleError(Exception):
    def __init__(self, code: int, what: str = ""):
        super().__init__(f"oracle error {code} {what}")
        self.code = code


def bloom_params(bit_count: int, hash_count: int, seed: int) -> BloomParams:
    p = BloomParams()
    rc = lib().orc_bloom_params_new(bit_count, hash_count, seed & (2**64 - 1), C.byref(p))
    if rc:
        raise OracleError(rc, "bloom_params_new")
    return p


def bloom_params_for_expected_items(n: int, fpr: float, seed: int) -> BloomParams:
    p = BloomParams()
    rc = lib().orc_bloom_params_for_expected_items(n, fpr, seed & (2**64 - 1), C.byref(p))
    if rc:
        raise OracleError(rc, "bloom_params_for_expected_items")
    return p


class Bloom:
    """AtomicBloomRef over a numpy u64 array (runtime_filter/src/bloom.rs:159-256)."""

    def __init__(self, params: BloomParams, words: Optional[np.ndarray] = None):
        self.p = params
        if words is None:
            words = np.zeros(params.word_count, dtype=np.uint64)
        rc = lib().orc_bloom_attach_check(C.byref(params), _ptr(words), words.size)
        if rc:
            raise OracleError(rc, "attach")
        self.words = words

    def clear(self):
        lib().orc_bloom_clear(C.byref(self.p), _ptr(self.words))

    def insert_u64(self, v: int):
        lib().orc_bloom_insert_hash(C.byref(self.p), _ptr(self.words), v & (2**64 - 1))

    def might_contain_u64(self, v: int) -> bool:
        return bool(lib().orc_bloom_might_contain_hash(C.byref(self.p), _ptr(self.words), v & (2**64 - 1)))

    def bit_index(self, v: int, i: int) -> int:
        return lib().orc_bloom_bit_index(C.byref(self.p), v & (2**64 - 1), i)

    def insert_keys(self, keys: np.ndarray, validity: Optional[np.ndarray] = None) -> int:
        keys = np.ascontiguousarray(keys)
        return lib().orc_bloom_insert_keys(C.byref(self.p), _ptr(self.words), _ptr(keys), keys.dtype.itemsize,
                                           _ptr(validity), keys.size)

    def probe_keys(self, keys: np.ndarray, validity: Optional[np.ndarray] = None):
        keys = np.ascontiguousarray(keys)
        keep = np.zeros(keys.size, dtype=np.uint8)
        rejected = lib().orc_bloom_probe_keys(C.byref(self.p), _ptr(self.words), _ptr(keys), keys.dtype.itemsize,
                                              _ptr(validity), keys.size, _ptr(keep))
        return keep, rejected


class Slot:
    """RuntimeFilterSlot lifecycle (runtime_filter/src/shared.rs:132-260)."""

    def __init__(self, params: BloomParams):
        self.bloom = Bloom(params)
        self.word = C.c_uint64(0)

    def snapshot(self):
        g, s = C.c_uint64(), C.c_int()
        lib().orc_lifecycle_unpack(self.word.value, C.byref(g), C.byref(s))
        return g.value, s.value

    def try_acquire_builder(self):
        g = C.c_uint64()
        rc = lib().orc_slot_try_acquire_builder(C.byref(self.word), C.byref(self.bloom.p), _ptr(self.bloom.words), C.byref(g))
        return rc, g.value

    def publish(self, gen):
        return lib().orc_slot_publish_build(C.byref(self.word), gen)

    def disable(self, gen):
        return lib().orc_slot_disable_build(C.byref(self.word), gen)

    def retire_ready(self, gen):
        return lib().orc_slot_retire_ready(C.byref(self.word), gen)

    def decision_for_u64(self, gen, v):
        return lib().orc_probe_decision_for_hash(C.byref(self.word), gen, C.byref(self.bloom.p), _ptr(self.bloom.words), v & (2**64 - 1))

    def decision_for_null(self, gen):
        return lib().orc_probe_decision_for_null(C.byref(self.word), gen)


# --------------------------------------------------------------------------- Layout

def specs(cols: Sequence[tuple]) -> "C.Array":
    arr = (ColumnSpec * max(1, len(cols)))()
    for i, (t, nullable) in enumerate(cols):
        arr[i].type_tag = t
        arr[i].nullable = 1 if nullable else 0
    return arr


def layout_plan(cols: Sequence[tuple], max_rows: int, block_size: int) -> LayoutPlan:
    plan = LayoutPlan()
    rc = lib().orc_layout_plan_new(specs(cols), len(cols), max_rows, block_size, C.byref(plan))
    if rc:
        raise OracleError(rc, "layout_plan_new")
    return plan


def fixed_row_cap(cols: Sequence[tuple], block_size: int) -> int:
    cap = C.c_uint32()
    rc = lib().orc_fixed_row_cap(specs(cols), len(cols), block_size, C.byref(cap))
    if rc:
        raise OracleError(rc, "fixed_row_cap")
    return cap.value


class Block:
    """BlockMut-style writer over a numpy byte buffer (page/arrow_layout/src/access.rs:236-636)."""

    def __init__(self, cols: Sequence[tuple], max_rows: int, block_size: int):
        self.cols = list(cols)
        self.plan = layout_plan(cols, max_rows, block_size)
        self.buf = np.zeros(block_size, dtype=np.uint8)
        rc = lib().orc_init_block(_ptr(self.buf), self.buf.size, C.byref(self.plan))
        if rc:
            raise OracleError(rc, "init_block")

    def _chk(self, rc, what):
        if rc:
            raise OracleError(rc, what)

    def write_fixed(self, col, row, data: bytes):
        b = np.frombuffer(data, dtype=np.uint8)
        self._chk(lib().orc_block_write_fixed(_ptr(self.buf), self.buf.size, col, row, _ptr(b), b.size), "write_fixed")

    def write_bool(self, col, row, v):
        self._chk(lib().orc_block_write_bool(_ptr(self.buf), self.buf.size, col, row, int(v)), "write_bool")

    def write_null(self, col, row):
        self._chk(lib().orc_block_write_null(_ptr(self.buf), self.buf.size, col, row), "write_null")

    def write_view_bytes(self, col, row, data: bytes):
        b = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(0, dtype=np.uint8)
        return lib().orc_block_write_view_bytes(_ptr(self.buf), self.buf.size, col, row, _ptr(b) if b.size else None, b.size)

    def commit_current_row(self):
        self._chk(lib().orc_block_commit_current_row(_ptr(self.buf), self.buf.size), "commit_current_row")

    def set_validity(self, col, row, valid):
        self._chk(lib().orc_block_set_validity(_ptr(self.buf), self.buf.size, col, row, int(valid)), "set_validity")

    def validate(self) -> int:
        return lib().orc_block_validate(_ptr(self.buf), self.buf.size)


def block_validate(buf: np.ndarray) -> int:
    return lib().orc_block_validate(_ptr(buf), buf.size)


def block_validate_v1(buf: np.ndarray) -> int:
    """BlockRef::open as the reference has it: type tags 1..9 only (no Decimal128 extension)."""
    return lib().orc_block_validate_v1(_ptr(buf), buf.size)


def import_check(kind: int, flags: int, buf: np.ndarray, cols: Sequence[tuple]) -> int:
    return lib().orc_import_check(kind, flags, _ptr(buf), buf.size, specs(cols), len(cols))


def page_header(kind: int, flags: int, payload_len: int) -> bytes:
    out = np.zeros(20, dtype=np.uint8)
    lib().orc_page_header_encode(kind, flags, payload_len, _ptr(out))
    return out.tobytes()


def page_header_decode(b: bytes):
    a = np.frombuffer(b, dtype=np.uint8)
    k, f, n = C.c_uint16(), C.c_uint16(), C.c_uint32()
    rc = lib().orc_page_header_decode(_ptr(a), C.byref(k), C.byref(f), C.byref(n))
    if rc:
        raise OracleError(rc, "page_header_decode")
    return k.value, f.value, n.value


# --------------------------------------------------------------------------- Operators

class OTable:
    """Decoded table (concatenated page columns)."""

    def __init__(self, t: Table, cols: Sequence[tuple]):
        self.t = t
        self.cols = list(cols)

    @staticmethod
    def from_pages(pages: np.ndarray, page_stride: int, cols: Sequence[tuple]) -> "OTable":
        pages = np.ascontiguousarray(pages).reshape(-1)
        npages = pages.size // page_stride
        t = Table()
        rc = lib().orc_table_from_pages(_ptr(pages), npages, page_stride, specs(cols), len(cols), C.byref(t))
        if rc:
            raise OracleError(rc, "table_from_pages")
        return OTable(t, cols)

    @property
    def rows(self) -> int:
        return self.t.rows

    def column(self, i: int):
        """Return a python list / numpy array of column i (None for nulls)."""
        c = self.t.cols[i]
        n = c.rows
        tag = c.type_tag
        valid = None
        if c.validity:
            valid = np.ctypeslib.as_array(C.cast(c.validity, C.POINTER(C.c_uint8)), shape=(n,)).copy() if n else np.zeros(0, np.uint8)
        if tag in (T_UTF8VIEW, T_BINARYVIEW):
            out = []
            raw = (C.c_uint8 * (16 * n)).from_address(c.values) if n else b""
            arr = np.frombuffer(raw, dtype=np.uint8).reshape(n, 16) if n else np.zeros((0, 16), np.uint8)
            for r in range(n):
                ln = int(arr[r, :4].view(np.uint32)[0])
                ptr = int(arr[r, 8:16].view(np.uint64)[0])
                s = C.string_at(ptr, ln) if ln else b""
                out.append(None if (valid is not None and not valid[r]) else s)
            return out
        dt = {T_BOOLEAN: np.uint8, T_INT16: np.int16, T_INT32: np.int32, T_INT64: np.int64,
              T_FLOAT32: np.float32, T_FLOAT64: np.float64}.get(tag)
        if dt is None:
            raw = np.ctypeslib.as_array(C.cast(c.values, C.POINTER(C.c_uint8)), shape=(n * 16,)).copy()
            return raw.reshape(n, 16), valid
        arr = np.ctypeslib.as_array(C.cast(c.values, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(n,)).copy() if n else np.zeros(0, dt)
        return arr, valid

    def select(self, keep: np.ndarray) -> "OTable":
        t = Table()
        keep = np.ascontiguousarray(keep, dtype=np.uint8)
        rc = lib().orc_table_select(C.byref(self.t), _ptr(keep), C.byref(t))
        if rc:
            raise OracleError(rc, "table_select")
        return OTable(t, self.cols)

    def take(self, rows: np.ndarray) -> "OTable":
        t = Table()
        rows = np.ascontiguousarray(rows, dtype=np.uint64)
        rc = lib().orc_table_take(C.byref(self.t), _ptr(rows), rows.size, C.byref(t))
        if rc:
            raise OracleError(rc, "table_take")
        return OTable(t, self.cols)

    def __del__(self):
        try:
            lib().orc_table_free(C.byref(self.t))
        except Exception:
            pass


class Expr:
    """Tiny expression builder producing postfix XNode lists."""

    def __init__(self, nodes: List[XNode]):
        self.nodes = nodes

    @staticmethod
    def col(i: int, source: int = 0) -> "Expr":
        n = XNode()
        n.op, n.a, n.b = X_COL, i, source
        return Expr([n])

    @staticmethod
    def f64(v: float) -> "Expr":
        n = XNode()
        n.op, n.f = X_LIT_F64, v
        return Expr([n])

    @staticmethod
    def i64(v: int) -> "Expr":
        n = XNode()
        n.op, n.i = X_LIT_I64, v
        return Expr([n])

    @staticmethod
    def i128(v: int) -> "Expr":
        n = XNode()
        u = v & (2**128 - 1)
        lo, hi = u & (2**64 - 1), u >> 64
        n.op = X_LIT_I128
        n.i = lo - 2**64 if lo >= 2**63 else lo
        n.i2 = hi - 2**64 if hi >= 2**63 else hi
        return Expr([n])

    @staticmethod
    def s(v: bytes) -> "Expr":
        n = XNode()
        assert len(v) <= 16
        n.op, n.slen, n.s = X_LIT_STR, len(v), v
        return Expr([n])

    def _bin(self, other: "Expr", op: int) -> "Expr":
        n = XNode()
        n.op = op
        return Expr(self.nodes + other.nodes + [n])

    def __add__(self, o): return self._bin(o, X_ADD)
    def __sub__(self, o): return self._bin(o, X_SUB)
    def __mul__(self, o): return self._bin(o, X_MUL)
    def lt(self, o): return self._bin(o, X_LT)
    def le(self, o): return self._bin(o, X_LE)
    def gt(self, o): return self._bin(o, X_GT)
    def ge(self, o): return self._bin(o, X_GE)
    def eq(self, o): return self._bin(o, X_EQ)
    def ne(self, o): return self._bin(o, X_NE)
    def and_(self, o): return self._bin(o, X_AND)


def _value_py(v: Value):
    if v.kind == V_NULL:
        return None
    if v.kind == V_F64:
        return v.f
    if v.kind == V_I64:
        return int(v.lo)
    if v.kind == V_I128:
        u = ((v.hi & (2**64 - 1)) << 64) | (v.lo & (2**64 - 1))
        return u - 2**128 if u >= 2**127 else u
    if v.kind == V_STR:
        return bytes(v.s[: v.slen]) if v.slen else b""
    if v.kind == V_BOOL:
        return bool(v.lo)
    raise ValueError(v.kind)


@dataclass
class AggOut:
    keys: list      # list of tuples
    aggs: list      # list of tuples
    rows_in: int
    rows_filtered: int
    rows_joined: int

    def by_key(self):
        return {k: a for k, a in zip(self.keys, self.aggs)}


def aggregate(scan: OTable, filt: Optional[Expr], keys: Sequence[Expr], aggs: Sequence[tuple],
              joins: Sequence[tuple] = (), sum_lanes: int = 0, batch_rows: int = 8192) -> AggOut:
    """aggs: sequence of (func, Expr|None).  joins: sequence of (OTable build, build_col, probe_src, probe_col)."""
    nodes: List[XNode] = []

    def add(e: Optional[Expr]):
        off = len(nodes)
        if e is None:
            return off, 0
        nodes.extend(e.nodes)
        return off, len(e.nodes)

    foff, flen = add(filt)
    koff, klen = [], []
    for k in keys:
        o, l = add(k)
        koff.append(o)
        klen.append(l)
    specs_ = (AggSpec * max(1, len(aggs)))()
    for j, (func, e) in enumerate(aggs):
        o, l = add(e)
        specs_[j].func, specs_[j].expr_off, specs_[j].expr_len = func, o, l
    narr = (XNode * max(1, len(nodes)))(*nodes)
    jarr = (JoinEdge * max(1, len(joins)))()
    for j, (bt, bcol, psrc, pcol) in enumerate(joins):
        jarr[j].build = C.pointer(bt.t)
        jarr[j].build_col, jarr[j].probe_src, jarr[j].probe_col = bcol, psrc, pcol
    ko = (C.c_int32 * max(1, len(keys)))(*koff)
    kl = (C.c_int32 * max(1, len(keys)))(*klen)
    res = AggResult()
    rc = lib().orc_aggregate(C.byref(scan.t), narr, foff, flen, jarr, len(joins), ko, kl, len(keys), specs_,
                             len(aggs), sum_lanes, batch_rows, C.byref(res))
    if rc:
        raise OracleError(rc, "aggregate")
    try:
        nk, na = res.nkeys, res.naggs
        keys_out = [tuple(_value_py(res.keys[g * nk + k]) for k in range(nk)) for g in range(res.ngroups)]
        aggs_out = [tuple(_value_py(res.aggs[g * na + j]) for j in range(na)) for g in range(res.ngroups)]
        return AggOut(keys_out, aggs_out, res.rows_in, res.rows_filtered, res.rows_joined)
    finally:
        lib().orc_agg_result_free(C.byref(res))


def filter_rows(scan: OTable, filt: Expr) -> np.ndarray:
    narr = (XNode * len(filt.nodes))(*filt.nodes)
    keep = np.zeros(scan.rows, dtype=np.uint8)
    kept = C.c_uint64()
    rc = lib().orc_filter(C.byref(scan.t), narr, 0, len(filt.nodes), _ptr(keep), C.byref(kept))
    if rc:
        raise OracleError(rc, "filter")
    return keep


def hash_join_pairs(build: OTable, build_col: int, probe: OTable, probe_col: int):
    n = C.c_uint64()
    rc = lib().orc_hash_join_pairs(C.byref(build.t), build_col, C.byref(probe.t), probe_col, None, None, 0, C.byref(n))
    if rc:
        raise OracleError(rc, "hash_join_pairs")
    b = np.zeros(n.value, dtype=np.uint64)
    p = np.zeros(n.value, dtype=np.uint64)
    rc = lib().orc_hash_join_pairs(C.byref(build.t), build_col, C.byref(probe.t), probe_col, _ptr(b), _ptr(p), n.value, C.byref(n))
    if rc:
        raise OracleError(rc, "hash_join_pairs")
    return b, p


def q6_pages(pages: np.ndarray, page_stride: int, nthreads: int, cols=(0, 1, 2, 3), date_lo=b"1994-01-01",
             date_hi=b"1995-01-01", disc_lo=0.05, disc_hi=0.07, qty_lt=24.0):
    pages = np.ascontiguousarray(pages).reshape(-1)
    res = Q6Result()
    carr = (C.c_int32 * 4)(*cols)
    rc = lib().orc_q6_pages(_ptr(pages), pages.size // page_stride, page_stride, nthreads, carr, date_lo, date_hi,
                            disc_lo, disc_hi, qty_lt, C.byref(res))
    if rc:
        raise OracleError(rc, "q6_pages")
    return res.sum, res.rows_in, res.rows_kept


def q1_pages(pages: np.ndarray, page_stride: int, nthreads: int, cols=(0, 1, 2, 3, 4, 5, 6),
             date_le=b"1998-09-02", with_tax=True, compensated=False):
    """compensated=True: Neumaier sums, i.e. the correctly rounded sums (full-size parity tests)."""
    pages = np.ascontiguousarray(pages).reshape(-1)
    res = Q1Result()
    carr = (C.c_int32 * 7)(*cols)
    rc = lib().orc_q1_pages(_ptr(pages), pages.size // page_stride, page_stride, nthreads, carr, date_le,
                            int(with_tax) | (2 if compensated else 0), C.byref(res))
    if rc:
        raise OracleError(rc, "q1_pages")
    out = {}
    for g in range(res.ngroups):
        gr = res.groups[g]
        out[(gr.returnflag, gr.linestatus)] = dict(sum_qty=gr.sum_qty, sum_base_price=gr.sum_base_price,
                                                   sum_disc_price=gr.sum_disc_price, sum_charge=gr.sum_charge,
                                                   sum_disc=gr.sum_disc, count=gr.count)
    return out, res.rows_in


class Q3Stream:
    """The TPC-H Q3 shape fed shard by shard (orc_q3.c): customer pages, then orders pages, then lineitem
    pages, each in any number of calls; SF100 never has to exist on the host at once."""

    def __init__(self, segment: bytes = b"BUILDING", date: bytes = b"1995-03-15", nthreads: int = 1):
        self.h = lib().orc_q3_new(segment, date)
        if not self.h:
            raise OracleError(-1, "orc_q3_new")
        self.nthreads = nthreads
        self._stage = 0

    def _pages(self, fn, pages, stride, what):
        pages = np.ascontiguousarray(pages).reshape(-1)
        rc = fn(self.h, _ptr(pages), pages.size // stride, stride, self.nthreads)
        if rc:
            raise OracleError(rc, what)

    def customer(self, pages, stride=65536):
        assert self._stage == 0
        self._pages(lib().orc_q3_customer, pages, stride, "q3 customer")

    def orders(self, pages, stride=65536):
        if self._stage == 0:
            if lib().orc_q3_customer_finish(self.h):
                raise OracleError(-1, "q3 customer_finish")
            self._stage = 1
        assert self._stage == 1
        self._pages(lib().orc_q3_orders, pages, stride, "q3 orders")

    def lineitem(self, pages, stride=65536):
        if self._stage == 0:
            self.orders(np.zeros(0, dtype=np.uint8))
        if self._stage == 1:
            if lib().orc_q3_orders_finish(self.h, self.nthreads):
                raise OracleError(-1, "q3 orders_finish")
            self._stage = 2
        self._pages(lib().orc_q3_lineitem, pages, stride, "q3 lineitem")

    def stats(self):
        out = (C.c_uint64 * 6)()
        lib().orc_q3_stats(self.h, out)
        return dict(customers=out[0], orders=out[1], lineitem_rows=out[2], filtered=out[3], joined=out[4], matched_orders=out[5])

    def groups(self):
        """{(l_orderkey, o_orderdate bytes, o_shippriority): (revenue, rows)}"""
        n = self.stats()["matched_orders"]
        keys = np.zeros(n, dtype=np.int32)
        dates = np.zeros((n, 12), dtype=np.uint8)
        lens = np.zeros(n, dtype=np.int32)
        prios = np.zeros(n, dtype=np.int32)
        sums = np.zeros(n, dtype=np.float64)
        cnts = np.zeros(n, dtype=np.uint64)
        got = lib().orc_q3_groups(self.h, _ptr(keys), _ptr(dates), _ptr(lens), _ptr(prios), _ptr(sums), _ptr(cnts), n)
        assert got == n
        out = {}
        for i in range(n):
            k = (int(keys[i]), bytes(dates[i, : lens[i]]), int(prios[i]))
            if k in out:   # build rows with equal (key, date, priority) are one group
                out[k] = (out[k][0] + float(sums[i]), out[k][1] + int(cnts[i]))
            else:
                out[k] = (float(sums[i]), int(cnts[i]))
        return out

    def close(self):
        if self.h:
            lib().orc_q3_free(self.h)
            self.h = None

    def __del__(self):
        self.close()
