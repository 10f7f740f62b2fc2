"""Host mirror of the worker-side hot path: device context, scans, the runtime Bloom filter
(same names as pg_fusion's `runtime_filter` crate) and fused pipelines (the operator chain
`worker_runtime` plans over a scan, worker_runtime/src/runtime.rs:667-698).

Everything here is plumbing over the C ABI of libpgf_b200.so; all compute runs in the CUDA
kernels.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from enum import IntEnum
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _lib
from .arrow_layout import ColumnSpec, TypeTag, _specs
from .errors import PgfError


class Cmp(IntEnum):
    LT = 0
    LE = 1
    GT = 2
    GE = 3
    EQ = 4
    NE = 5


class AggFunc(IntEnum):
    SUM = 1
    AVG = 2
    COUNT_STAR = 3
    COUNT = 4


class RuntimeFilterState(IntEnum):  # runtime_filter/src/shared.rs:11-23
    Free = 0
    Building = 1
    Ready = 2
    Disabled = 3


class ProbeDecision(IntEnum):  # runtime_filter/src/shared.rs:46-55
    PassUnfiltered = 0
    MaybePresent = 1
    DefinitelyAbsent = 2


class GenTable(IntEnum):
    LINEITEM_Q6 = 1
    LINEITEM_Q1 = 2
    LINEITEM_Q3 = 3
    ORDERS_Q3 = 4
    CUSTOMER_Q3 = 5
    KEYS_I64 = 6
    LINEITEM_Q6_D = 7   # Decimal128 money, Date32 dates (SURVEY 8d "D" variant)
    LINEITEM_Q1_D = 8


@dataclass(frozen=True)
class BloomParams:
    """BloomParams (runtime_filter/src/bloom.rs:17-100)."""
    bit_count: int
    word_count: int
    hash_count: int
    seed: int

    @staticmethod
    def new(bit_count: int, hash_count: int, seed: int) -> "BloomParams":
        p = _lib.BloomParamsC()
        rc = _lib.lib().pgf_bloom_params_new(bit_count, hash_count, seed & (2**64 - 1), C.byref(p))
        if rc:
            raise PgfError(rc, "BloomParams::new")
        return BloomParams(p.bit_count, p.word_count, p.hash_count, p.seed)

    @staticmethod
    def for_expected_items(expected_items: int, false_positive_rate: float, seed: int) -> "BloomParams":
        p = _lib.BloomParamsC()
        rc = _lib.lib().pgf_bloom_params_for_expected_items(expected_items, false_positive_rate,
                                                            seed & (2**64 - 1), C.byref(p))
        if rc:
            raise PgfError(rc, "BloomParams::for_expected_items")
        return BloomParams(p.bit_count, p.word_count, p.hash_count, p.seed)

    def _c(self) -> _lib.BloomParamsC:
        return _lib.BloomParamsC(self.bit_count, self.word_count, self.hash_count, self.seed)


# GUC defaults: pg/extension/src/guc.rs:41-46 (bits, hashes, seed "pgfusion")
GUC_DEFAULT_BLOOM = dict(bit_count=1 << 20, hash_count=4, seed=0x7067667573696F6E)


def _value_py(v: _lib.Value):
    if v.kind == 0:
        return None
    if v.kind == 1:
        return v.f64
    if v.kind == 2:
        return int(v.lo)
    if v.kind == 3:
        u = ((v.hi & (2**64 - 1)) << 64) | (v.lo & (2**64 - 1))
        return u - 2**128 if u >= 2**127 else u
    if v.kind == 4:
        return bytes(v.str[: v.slen])
    raise ValueError(f"bad value kind {v.kind}")


def encode_result_pages(res, page_size: int = 65536):
    """pgf_result -> (transport schema, pages) through pgf_result_schema / pgf_result_encode_pages."""
    L = _lib.lib()
    arr = (_lib.ColumnSpec * 20)()
    n = C.c_uint32()
    rc = L.pgf_result_schema(res, arr, C.byref(n))
    if rc:
        raise PgfError(rc, "pgf_result_schema")
    schema = [ColumnSpec(TypeTag(arr[i].type_tag), bool(arr[i].nullable)) for i in range(n.value)]
    cap = C.c_uint32()
    rc = L.pgf_layout_fixed_row_cap(arr, n.value, page_size - 20, C.byref(cap))
    if rc or cap.value == 0:
        raise PgfError(rc or 113, "result row does not fit a page")
    ngroups = res.contents.ngroups
    npages = (ngroups + cap.value - 1) // cap.value
    pages = np.zeros((max(npages, 1), page_size), dtype=np.uint8)
    got, rows = C.c_uint64(), C.c_uint64()
    rc = L.pgf_result_encode_pages(res, page_size, 0, pages.ctypes.data_as(C.c_void_p), npages, C.byref(got), C.byref(rows))
    if rc:
        raise PgfError(rc, "pgf_result_encode_pages")
    assert got.value == npages and rows.value == ngroups
    return schema, pages[:npages]


@dataclass
class PipelineResult:
    rows_in: int
    rows_bloom: int
    rows_filtered: int
    rows_out: int
    keys: List[tuple]
    aggs: List[tuple]
    join_table: int
    bloom_rows: int
    kernel_ms: float
    kernel_launches: int
    variant: str = ""                                  # kernel instantiation that ran (registered shape or "generic")
    result_schema: Optional[List[ColumnSpec]] = None   # transport schema of the result pages
    result_pages: Optional[np.ndarray] = None          # [npages, page_size] uint8, reference page format

    def by_key(self) -> Dict[tuple, tuple]:
        return dict(zip(self.keys, self.aggs))


def _literal(v, type_tag: Optional[int] = None) -> _lib.Literal:
    lit = _lib.Literal()
    if isinstance(v, bool):   # (before int: bool is an int in Python)
        lit.type_tag, lit.i64 = int(TypeTag.Boolean), int(v)
    elif isinstance(v, float):
        lit.type_tag, lit.f64 = int(TypeTag.Float64), v
    elif isinstance(v, int):
        if type_tag == int(TypeTag.Decimal128) or not (-(2**63) <= v < 2**63):
            u = v & (2**128 - 1)
            lo, hi = u & (2**64 - 1), u >> 64
            lit.type_tag = int(TypeTag.Decimal128)
            lit.i64 = lo - 2**64 if lo >= 2**63 else lo
            lit.hi = hi - 2**64 if hi >= 2**63 else hi
        else:
            lit.type_tag, lit.i64 = int(TypeTag.Int64), v
            lit.hi = -1 if v < 0 else 0
    elif isinstance(v, (bytes, str)):
        b = v.encode() if isinstance(v, str) else v
        if len(b) > 16:
            raise ValueError("string literal too long")
        lit.type_tag, lit.slen = int(TypeTag.Utf8View), len(b)
        for i, ch in enumerate(b):
            lit.str[i] = ch
    else:
        raise TypeError(f"unsupported literal {v!r}")
    return lit


ColRefLike = Union[int, Tuple[int, int]]


def _colref(c: ColRefLike) -> _lib.ColRef:
    if isinstance(c, tuple):
        return _lib.ColRef(c[0], c[1])
    return _lib.ColRef(0, int(c))


@dataclass
class Factor:
    """One factor of a projection expression: x, (c - x) or (c + x)."""
    col: ColRefLike
    kind: int = 0  # 0 = x, 1 = c - x, 2 = c + x
    c: Union[int, float] = 0

    @staticmethod
    def of(col: ColRefLike) -> "Factor":
        return Factor(col, 0, 0)

    @staticmethod
    def const_minus(c, col: ColRefLike) -> "Factor":
        return Factor(col, 1, c)

    @staticmethod
    def const_plus(c, col: ColRefLike) -> "Factor":
        return Factor(col, 2, c)


class PipelineBuilder:
    """Builds the POD plan (pgf_pipeline) of one fused pipeline over a scan."""

    def __init__(self, ctx: "Context", scan: "Scan"):
        self.ctx = ctx
        self.p = _lib.Pipeline()
        self.p.scan_id = scan.scan_id
        self.scan = scan

    def bloom_probe(self, rf: "RuntimeFilter", key: ColRefLike, generation: Optional[int] = None) -> "PipelineBuilder":
        i = self.p.nbloom
        self.p.bloom[i].bloom = rf.handle
        self.p.bloom[i].expected_generation = rf.generation if generation is None else generation
        self.p.bloom[i].key = _colref(key)
        self.p.nbloom = i + 1
        return self

    def filter(self, col: ColRefLike, cmp: Cmp, literal) -> "PipelineBuilder":
        i = self.p.nterms
        if i >= _lib.MAX_TERMS:
            raise PgfError(6, "too many predicate terms")
        self.p.terms[i].col = _colref(col)
        self.p.terms[i].cmp = int(cmp)
        tt = None
        if not isinstance(col, tuple):
            tt = int(self.scan.schema[col].type_tag)
        self.p.terms[i].lit = _literal(literal, tt)
        self.p.nterms = i + 1
        return self

    def join(self, join_table: int, probe_key: ColRefLike) -> "PipelineBuilder":
        i = self.p.njoins
        self.p.joins[i].join_table = join_table
        self.p.joins[i].probe_key = _colref(probe_key)
        self.p.njoins = i + 1
        return self

    def _expr(self, factors: Sequence[Factor]) -> int:
        e = self.p.nexprs
        if e >= _lib.MAX_EXPRS:
            raise PgfError(6, "too many expressions")
        self.p.exprs[e].nfactors = len(factors)
        for k, f in enumerate(factors):
            self.p.exprs[e].factors[k].kind = f.kind
            self.p.exprs[e].factors[k].col = _colref(f.col)
            tt = None
            if not isinstance(f.col, tuple):
                tt = int(self.scan.schema[f.col].type_tag)
            self.p.exprs[e].factors[k].c = _literal(f.c, tt)
        self.p.nexprs = e + 1
        return e

    def aggregate(self, keys: Sequence[ColRefLike], aggs: Sequence[Tuple[AggFunc, Optional[Sequence[Factor]]]],
                  expected_groups: int = 0) -> "PipelineBuilder":
        """aggs: (func, factors) where factors is the product expression (None for COUNT(*)).
        Identical expressions share one accumulator (SUM(x) and AVG(x) share sum + count)."""
        self.p.sink = 1
        self.p.nkeys = len(keys)
        for i, k in enumerate(keys):
            self.p.keys[i] = _colref(k)
        seen: Dict[tuple, int] = {}
        self.p.naggs = len(aggs)
        for j, (func, factors) in enumerate(aggs):
            self.p.aggs[j].func = int(func)
            if func == AggFunc.COUNT_STAR or factors is None:
                self.p.aggs[j].expr = -1
                continue
            sig = tuple((f.col, f.kind, f.c) for f in factors)
            if sig not in seen:
                seen[sig] = self._expr(factors)
            self.p.aggs[j].expr = seen[sig]
        self.p.expected_groups = expected_groups
        return self

    def order_by(self, terms: Sequence[tuple], limit: int = 0) -> "PipelineBuilder":
        """SortExec / TopK above the aggregate.  terms: ("key"|"agg", index, descending[, nulls_first]);
        NULL order defaults to DataFusion's (ASC NULLS LAST, DESC NULLS FIRST)."""
        self.p.nsort = len(terms)
        for i, t in enumerate(terms):
            which, index, desc = t[0], t[1], bool(t[2])
            nulls_first = bool(t[3]) if len(t) > 3 else desc
            self.p.sort[i] = _lib.SortKey(1 if which == "agg" else 0, index, int(desc), int(nulls_first))
        self.p.limit = limit
        return self

    def build_join(self, key: ColRefLike, payload: Sequence[ColRefLike] = (),
                   bloom: Optional["RuntimeFilter"] = None, expected_rows: int = 0, rows_only: bool = False) -> "PipelineBuilder":
        """expected_rows: sizing hint for the build-row buffer (0: as many rows as the scan holds); a hint that
        turns out too small costs one more pass, never a wrong result."""
        self.p.sink = 2
        self.p.expected_groups = expected_rows
        self.p.build_flags = 1 if rows_only else 0
        self.p.build_key = _colref(key)
        self.p.npayload = len(payload)
        for i, c in enumerate(payload):
            self.p.payload[i] = _colref(c)
        self.p.build_bloom = bloom.handle if bloom is not None else 0
        return self

    def count(self) -> "PipelineBuilder":
        self.p.sink = 3
        return self

    def check(self) -> int:
        return _lib.lib().pgf_pipeline_check(self.ctx.h, C.byref(self.p))

    def run(self, pages: bool = False) -> PipelineResult:
        """pages=True also encodes the output rows as reference result pages (ResultPageProducer)."""
        res = C.POINTER(_lib.Result)()
        self.ctx._check(_lib.lib().pgf_pipeline_run(self.ctx.h, C.byref(self.p), C.byref(res)))
        return self.ctx._take_result(res, pages)

    def run_partial(self, dev_ptr: int, capacity_bytes: int) -> Tuple[int, PipelineResult]:
        res = C.POINTER(_lib.Result)()
        nbytes = C.c_uint64()
        self.ctx._check(_lib.lib().pgf_pipeline_run_partial(self.ctx.h, C.byref(self.p), dev_ptr, capacity_bytes,
                                                            C.byref(nbytes), C.byref(res)))
        return nbytes.value, self.ctx._take_result(res)

    def partial_state_bytes(self, max_groups: int) -> int:
        n = C.c_uint64()
        self.ctx._check(_lib.lib().pgf_partial_state_bytes(C.byref(self.p), max_groups, C.byref(n)))
        return n.value

    def run_sharded(self, max_groups: int = 1) -> PipelineResult:
        """AggregateExec Partial -> all-gather -> Final inside the library (pgf_pipeline_run_sharded)."""
        res = C.POINTER(_lib.Result)()
        self.ctx._check(_lib.lib().pgf_pipeline_run_sharded(self.ctx.h, C.byref(self.p), max_groups, C.byref(res)))
        return self.ctx._take_result(res)

    def run_partial_async(self, dev_ptr: int, capacity_bytes: int) -> None:
        """Enqueue kernel + partial-state extraction on the compute stream; no synchronisation."""
        self.ctx._check(_lib.lib().pgf_pipeline_run_partial_async(self.ctx.h, C.byref(self.p), dev_ptr, capacity_bytes))

    def merge_partials_bounded(self, dev_ptr: int, stride_bytes: int, nstates: int) -> PipelineResult:
        res = C.POINTER(_lib.Result)()
        self.ctx._check(_lib.lib().pgf_pipeline_merge_partials_bounded(self.ctx.h, C.byref(self.p), dev_ptr, stride_bytes,
                                                                       nstates, C.byref(res)))
        return self.ctx._take_result(res)

    def merge_partials(self, dev_ptr: int, stride_bytes: int, nstates: int) -> PipelineResult:
        res = C.POINTER(_lib.Result)()
        self.ctx._check(_lib.lib().pgf_pipeline_merge_partials(self.ctx.h, C.byref(self.p), dev_ptr, stride_bytes,
                                                               nstates, C.byref(res)))
        return self.ctx._take_result(res)


class Scan:
    """A declared scan: pages pushed here live in HBM (WorkerPgScanExec + ArrowPageDecoder)."""

    def __init__(self, ctx: "Context", scan_id: int, schema: Sequence[ColumnSpec]):
        self.ctx = ctx
        self.scan_id = scan_id
        self.schema = list(schema)

    def push_page(self, page: np.ndarray) -> None:
        page = np.ascontiguousarray(page, dtype=np.uint8)
        self.ctx._check(_lib.lib().pgf_scan_push_page(self.ctx.h, self.scan_id, page.ctypes.data_as(C.c_void_p), page.size))

    def push_pages(self, pages: np.ndarray, stride: Optional[int] = None) -> None:
        pages = np.ascontiguousarray(pages, dtype=np.uint8)
        if stride is None:
            stride = pages.shape[-1] if pages.ndim == 2 else self.ctx.page_size
        n = pages.size // stride
        self.ctx._check(_lib.lib().pgf_scan_push_pages(self.ctx.h, self.scan_id, pages.ctypes.data_as(C.c_void_p), n, stride))

    def push_pages_ptr(self, ptr: int, npages: int, stride: int) -> None:
        """Raw host pointer (e.g. a pinned torch tensor's data_ptr())."""
        self.ctx._check(_lib.lib().pgf_scan_push_pages(self.ctx.h, self.scan_id, ptr, npages, stride))

    def finish(self) -> None:
        self.ctx._check(_lib.lib().pgf_scan_finish(self.ctx.h, self.scan_id))

    def info(self) -> _lib.ScanInfo:
        out = _lib.ScanInfo()
        self.ctx._check(_lib.lib().pgf_scan_get_info(self.ctx.h, self.scan_id, C.byref(out)))
        return out

    def reset(self) -> None:
        self.ctx._check(_lib.lib().pgf_scan_reset(self.ctx.h, self.scan_id))

    def release(self) -> None:
        self.ctx._check(_lib.lib().pgf_scan_release(self.ctx.h, self.scan_id))

    def read_pages(self, first: int = 0, n: Optional[int] = None) -> np.ndarray:
        info = self.info()
        n = info.pages - first if n is None else n
        out = np.zeros((n, self.ctx.page_size), dtype=np.uint8)
        self.ctx._check(_lib.lib().pgf_scan_read_pages(self.ctx.h, self.scan_id, first, n, out.ctypes.data_as(C.c_void_p)))
        return out

    def pipeline(self) -> PipelineBuilder:
        return PipelineBuilder(self.ctx, self)


class RuntimeFilter:
    """RuntimeFilterSlot + builder/probe handles over a Bloom bit array in HBM
    (runtime_filter/src/shared.rs:132-374)."""

    def __init__(self, ctx: "Context", params: BloomParams):
        self.ctx = ctx
        self.params = params
        h = C.c_uint64()
        ctx._check(_lib.lib().pgf_bloom_create(ctx.h, C.byref(params._c()), C.byref(h)))
        self.handle = h.value
        self.generation = 0

    def snapshot(self) -> Tuple[int, RuntimeFilterState]:
        g, s = C.c_uint64(), C.c_int32()
        self.ctx._check(_lib.lib().pgf_bloom_snapshot(self.ctx.h, self.handle, C.byref(g), C.byref(s)))
        return g.value, RuntimeFilterState(s.value)

    def try_acquire_builder(self) -> int:
        g = C.c_uint64()
        self.ctx._check(_lib.lib().pgf_bloom_begin_build(self.ctx.h, self.handle, C.byref(g)))
        self.generation = g.value
        return g.value

    def insert_keys(self, keys: np.ndarray, validity: Optional[np.ndarray] = None) -> int:
        keys = np.ascontiguousarray(keys)
        n = C.c_uint64()
        vb = None if validity is None else np.packbits(np.asarray(validity, dtype=bool), bitorder="little")
        self.ctx._check(_lib.lib().pgf_bloom_insert_keys(
            self.ctx.h, self.handle, keys.ctypes.data_as(C.c_void_p), keys.dtype.itemsize,
            None if vb is None else vb.ctypes.data_as(C.c_void_p), keys.size, C.byref(n)))
        return n.value

    def insert_u64(self, v: int) -> None:
        self.insert_keys(np.array([v & (2**64 - 1)], dtype=np.uint64).view(np.int64))

    def insert_scan(self, scan: Scan, col: int) -> int:
        n = C.c_uint64()
        self.ctx._check(_lib.lib().pgf_bloom_insert_scan(self.ctx.h, self.handle, scan.scan_id, col, C.byref(n)))
        return n.value

    def publish_ready(self) -> None:
        self.ctx._check(_lib.lib().pgf_bloom_publish_ready(self.ctx.h, self.handle))

    def disable(self) -> None:
        self.ctx._check(_lib.lib().pgf_bloom_disable_build(self.ctx.h, self.handle))

    def retire_ready_after_quiescence(self) -> None:
        self.ctx._check(_lib.lib().pgf_bloom_retire_ready(self.ctx.h, self.handle))

    def words(self) -> np.ndarray:
        out = np.zeros(self.params.word_count, dtype=np.uint64)
        self.ctx._check(_lib.lib().pgf_bloom_read_words(self.ctx.h, self.handle, out.ctypes.data_as(C.c_void_p), out.size))
        return out

    def or_words(self, words: np.ndarray) -> None:
        words = np.ascontiguousarray(words, dtype=np.uint64)
        self.ctx._check(_lib.lib().pgf_bloom_or_words(self.ctx.h, self.handle, words.ctypes.data_as(C.c_void_p), words.size))

    def publish_to_pool(self, base_ptr: int, length: int, slot_count: int, slot_index: int, generation: int) -> None:
        """Copy the GPU-built words into a slot of the shared-memory pool and publish it Ready
        (runtime_filter/src/pool.rs: insert + publish_ready), so backends probe it unchanged."""
        self.ctx._check(_lib.lib().pgf_bloom_publish_to_pool(self.ctx.h, self.handle, base_ptr, length, slot_count, slot_index, generation))

    def read_words_into(self, ptr: int) -> None:
        """Copy the word array into a caller buffer (host or device pointer) of word_count * 8 bytes."""
        self.ctx._check(_lib.lib().pgf_bloom_read_words(self.ctx.h, self.handle, C.c_void_p(ptr), self.params.word_count))

    def device_words_ptr(self) -> int:
        return _lib.lib().pgf_bloom_device_words(self.ctx.h, self.handle)

    def or_all_reduce(self) -> None:
        """Bloom OR-merge over the context's communicator (pgf_bloom_or_all_reduce)."""
        self.ctx._check(_lib.lib().pgf_bloom_or_all_reduce(self.ctx.h, self.handle))

    def or_device_words(self, dev_ptr: int, narrays: int) -> None:
        self.ctx._check(_lib.lib().pgf_bloom_or_device_words(self.ctx.h, self.handle, dev_ptr, self.params.word_count, narrays))

    def probe_keys(self, keys: np.ndarray, validity: Optional[np.ndarray] = None, generation: Optional[int] = None):
        keys = np.ascontiguousarray(keys)
        out = np.zeros(keys.size, dtype=np.uint8)
        st = _lib.ProbeStats()
        vb = None if validity is None else np.packbits(np.asarray(validity, dtype=bool), bitorder="little")
        self.ctx._check(_lib.lib().pgf_bloom_probe_keys(
            self.ctx.h, self.handle, self.generation if generation is None else generation,
            keys.ctypes.data_as(C.c_void_p), keys.dtype.itemsize, None if vb is None else vb.ctypes.data_as(C.c_void_p),
            keys.size, out.ctypes.data_as(C.c_void_p), C.byref(st)))
        return out, st

    def decision_for_u64(self, v: int, generation: Optional[int] = None) -> ProbeDecision:
        d, _ = self.probe_keys(np.array([v & (2**64 - 1)], dtype=np.uint64).view(np.int64), generation=generation)
        return ProbeDecision(int(d[0]))

    def decision_for_null(self, generation: Optional[int] = None) -> ProbeDecision:
        d, _ = self.probe_keys(np.zeros(1, dtype=np.int64), validity=np.array([False]), generation=generation)
        return ProbeDecision(int(d[0]))

    def probe_scan(self, scan: Scan, col: int, generation: Optional[int] = None):
        out = np.zeros(scan.info().rows, dtype=np.uint8)
        st = _lib.ProbeStats()
        self.ctx._check(_lib.lib().pgf_bloom_probe_scan(
            self.ctx.h, self.handle, self.generation if generation is None else generation, scan.scan_id, col,
            out.ctypes.data_as(C.c_void_p), C.byref(st)))
        return out, st

    def destroy(self) -> None:
        self.ctx._check(_lib.lib().pgf_bloom_destroy(self.ctx.h, self.handle))


class Context:
    """One device context (one per GPU process)."""

    def __init__(self, device: int = 0, page_size: int = 65536, staging_pages: int = 0,
                 keep_redundant_bloom_probes: bool = False):
        """keep_redundant_bloom_probes: PGF_CFG_KEEP_REDUNDANT_BLOOM_PROBES (by default a fused Bloom probe
        is dropped when the same pipeline probes a join table on that key, or when the filter is saturated)."""
        self.h = C.c_void_p()
        cfg = _lib.Config(device, page_size, staging_pages, 1 if keep_redundant_bloom_probes else 0)
        rc = _lib.lib().pgf_ctx_create(C.byref(cfg), C.byref(self.h))
        if rc:
            self.h = None
            raise PgfError(rc, "pgf_ctx_create (a CUDA device is required; there is no CPU fallback)")
        self.page_size = page_size
        self.device = device
        self._next_scan = 1

    def _check(self, rc: int) -> None:
        if rc:
            raise PgfError(rc, (_lib.lib().pgf_last_error(self.h) or b"").decode(errors="replace"))

    def _take_result(self, res, pages: bool = False) -> PipelineResult:
        try:
            r = res.contents
            nk, na = r.nkeys, r.naggs
            keys = [tuple(_value_py(r.keys[g * nk + k]) for k in range(nk)) for g in range(r.ngroups)]
            aggs = [tuple(_value_py(r.aggs[g * na + a]) for a in range(na)) for g in range(r.ngroups)]
            out = PipelineResult(r.rows_in, r.rows_bloom, r.rows_filtered, r.rows_out, keys, aggs, r.join_table,
                                 r.bloom_rows, r.kernel_ms, r.kernel_launches)
            out.variant = bytes(r.variant).split(b"\0")[0].decode()
            if pages and (nk or na):
                out.result_schema, out.result_pages = encode_result_pages(res, self.page_size)
            return out
        finally:
            _lib.lib().pgf_result_free(res)

    def declare_scan(self, schema: Sequence[ColumnSpec], expected_pages: int = 0, scan_id: Optional[int] = None) -> Scan:
        if scan_id is None:
            scan_id = self._next_scan
            self._next_scan += 1
        self._check(_lib.lib().pgf_scan_declare(self.h, scan_id, _specs(schema), len(schema), expected_pages))
        return Scan(self, scan_id, schema)

    def gen_scan(self, table: GenTable, rows: int, seed: int = 42, first_row: int = 0, scale_rows: int = 0,
                 dense_keys: bool = False, scan_id: Optional[int] = None) -> Scan:
        if scan_id is None:
            scan_id = self._next_scan
            self._next_scan += 1
        spec = _lib.GenSpec(int(table), int(dense_keys), seed, first_row, rows, scale_rows)
        self._check(_lib.lib().pgf_gen_scan(self.h, scan_id, C.byref(spec)))
        arr = (_lib.ColumnSpec * _lib.MAX_COLS)()
        n = C.c_uint32()
        self._check(_lib.lib().pgf_gen_schema(int(table), arr, C.byref(n)))
        schema = [ColumnSpec(TypeTag(arr[i].type_tag), bool(arr[i].nullable)) for i in range(n.value)]
        return Scan(self, scan_id, schema)

    def runtime_filter(self, params: BloomParams) -> RuntimeFilter:
        return RuntimeFilter(self, params)

    def register_host_region(self, ptr: int, nbytes: int) -> None:
        self._check(_lib.lib().pgf_ctx_register_host_region(self.h, ptr, nbytes))

    def unregister_host_region(self, ptr: int) -> None:
        self._check(_lib.lib().pgf_ctx_unregister_host_region(self.h, ptr))

    def join_table_info(self, handle: int):
        info = _lib.JoinInfo()
        self._check(_lib.lib().pgf_join_table_get_info(self.h, handle, C.byref(info)))
        return info

    def join_table_export(self, handle: int, dev_ptr: int, capacity_rows: int) -> int:
        """Occupied slots of a built table as self-contained records in a device buffer; returns the row count."""
        n = C.c_uint64()
        self._check(_lib.lib().pgf_join_table_export(self.h, handle, dev_ptr, capacity_rows, C.byref(n)))
        return n.value

    def join_table_from_fragments(self, like: int, dev_ptr: int, stride_bytes: int, counts: Sequence[int]) -> int:
        """Rebuild one table from the exported fragments of several ranks (broadcast join)."""
        arr = (C.c_uint64 * len(counts))(*[int(c) for c in counts])
        out = C.c_uint64()
        self._check(_lib.lib().pgf_join_table_from_fragments(self.h, like, dev_ptr, stride_bytes, arr, len(counts), C.byref(out)))
        return out.value

    def destroy_join_table(self, handle: int) -> None:
        self._check(_lib.lib().pgf_join_table_destroy(self.h, handle))

    # ---- multi-GPU: one process per GPU, the collectives run inside the library (NCCL over NVLink) ----
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        rc = _lib.lib().pgf_comm_unique_id(buf)
        if rc:
            raise PgfError(rc, "pgf_comm_unique_id")
        return bytes(buf)

    def comm_init(self, unique_id: bytes, rank: int, world: int) -> None:
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._check(_lib.lib().pgf_comm_init(self.h, buf, rank, world))

    def comm_info(self) -> Tuple[int, int]:
        r, w = C.c_int32(), C.c_int32()
        self._check(_lib.lib().pgf_comm_info(self.h, C.byref(r), C.byref(w)))
        return r.value, w.value

    def comm_destroy(self) -> None:
        self._check(_lib.lib().pgf_comm_destroy(self.h))

    def comm_all_gather(self, send_ptr: int, recv_ptr: int, nbytes: int) -> None:
        self._check(_lib.lib().pgf_comm_all_gather(self.h, send_ptr, recv_ptr, nbytes))

    def comm_all_gather_host(self, payload: bytes) -> List[bytes]:
        """All-gather of one equally sized host buffer per rank; returns the buffers in rank order."""
        _, world = self.comm_info()
        send = (C.c_uint8 * max(1, len(payload))).from_buffer_copy(payload or b"\0")
        recv = (C.c_uint8 * (max(1, len(payload)) * world))()
        self._check(_lib.lib().pgf_comm_all_gather_host(self.h, send, recv, len(payload)))
        raw = bytes(recv)
        return [raw[i * len(payload):(i + 1) * len(payload)] for i in range(world)]

    def exchange(self, handle: int, partition: bool, rows_only: bool = False, emulate: Optional[Tuple[int, int]] = None) -> Tuple[int, int]:
        """pgf_join_table_exchange: (new handle, bytes this rank sent over NVLink).  emulate=(world, rank): a context
        without a communicator plays one rank of a partition of its own rows (single-GPU tests)."""
        out, sent = C.c_uint64(), C.c_uint64()
        mode = (1 if partition else 0) | (4 if rows_only else 0)
        if emulate is not None:
            mode |= (emulate[0] & 0xFF) << 8 | (emulate[1] & 0xFF) << 16
        self._check(_lib.lib().pgf_join_table_exchange(self.h, handle, mode, C.byref(out), C.byref(sent)))
        return out.value, sent.value

    def row_set_pipeline(self, row_set: int, schema) -> "PipelineBuilder":
        """A pipeline that scans a row set (column 0 = key, column i + 1 = payload i) instead of pages."""
        class _RowScan:
            scan_id = 0
        rs = _RowScan()
        rs.schema = schema
        pb = PipelineBuilder(self, rs)
        pb.p.scan_row_set = row_set
        return pb

    def synchronize(self) -> None:
        self._check(_lib.lib().pgf_ctx_synchronize(self.h))

    def last_kernel_ms(self) -> float:
        return float(_lib.lib().pgf_ctx_last_kernel_ms(self.h))

    def runtime_filter_metrics(self) -> Dict[str, int]:
        """The RuntimeFilter* counters of this context (runtime_metrics/src/lib.rs:125-131), by field name."""
        m = _lib.RuntimeFilterMetrics()
        self._check(_lib.lib().pgf_ctx_runtime_filter_metrics(self.h, C.byref(m)))
        return {name: int(getattr(m, name)) for name, _ in m._fields_}

    def note_pool_exhausted(self) -> None:
        self._check(_lib.lib().pgf_ctx_note_pool_exhausted(self.h))

    def compute_stream(self) -> int:
        return _lib.lib().pgf_ctx_compute_stream(self.h)

    def close(self) -> None:
        if self.h:
            _lib.lib().pgf_ctx_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def device_count() -> int:
    n = C.c_int32()
    _lib.lib().pgf_device_count(C.byref(n))
    return n.value
