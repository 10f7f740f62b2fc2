"""ctypes declarations for libpgf_b200.so (include/pgf_b200.h).

The product path has no CPU fallback: if the shared library is missing this module raises
at load time, and contexts cannot be created without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PGF_B200_LIB") or os.path.join(_PKG_DIR, "libpgf_b200.so")  # override: kernel A/B experiments
CSRC_DIR = os.path.join(_PKG_DIR, "csrc")

MAX_COLS = 16
MAX_TERMS, MAX_JOINS, MAX_BLOOM, MAX_KEYS, MAX_EXPRS, MAX_AGGS, MAX_PAYLOAD = 8, 2, 2, 4, 8, 16, 4

u8, u16, u32, u64 = C.c_uint8, C.c_uint16, C.c_uint32, C.c_uint64
i32, i64 = C.c_int32, C.c_int64
vp = C.c_void_p
P = C.POINTER


class Config(C.Structure):
    _fields_ = [("device", i32), ("page_size", u32), ("staging_pages", u32), ("flags", u32)]


class ColumnSpec(C.Structure):
    _fields_ = [("type_tag", u16), ("nullable", u16)]


class ColumnLayout(C.Structure):
    _fields_ = [("type_tag", u16), ("flags", u16), ("validity_off", u32), ("values_off", u32),
                ("validity_len", u32), ("values_len", u32)]


class LayoutPlanC(C.Structure):
    _fields_ = [("block_size", u32), ("max_rows", u32), ("front_base", u32), ("pool_base", u32),
                ("ncols", u32), ("cols", ColumnLayout * 64)]


class ScanInfo(C.Structure):
    _fields_ = [("pages", u64), ("rows", u64), ("bytes", u64), ("ncols", u32), ("finished", u32)]


class BloomParamsC(C.Structure):
    _fields_ = [("bit_count", u64), ("word_count", u64), ("hash_count", u64), ("seed", u64)]


class ProbeStats(C.Structure):
    _fields_ = [("probe_rows", u64), ("rejected_rows", u64), ("pass_unfiltered", u64)]


class RuntimeFilterMetrics(C.Structure):
    """pgf_runtime_filter_metrics: the RuntimeFilter* counters of runtime_metrics/src/lib.rs:125-131."""
    _fields_ = [("allocated_total", u64), ("ready_total", u64), ("pool_exhausted_total", u64), ("build_rows_total", u64),
                ("probe_rows_total", u64), ("probe_rows_rejected_total", u64), ("probe_pass_unfiltered_total", u64)]


class Literal(C.Structure):
    _fields_ = [("type_tag", i32), ("slen", i32), ("f64", C.c_double), ("i64", i64), ("hi", i64), ("str", u8 * 16)]


class ColRef(C.Structure):
    _fields_ = [("source", i32), ("col", i32)]


class PredTerm(C.Structure):
    _fields_ = [("col", ColRef), ("cmp", i32), ("reserved", i32), ("lit", Literal)]


class Factor(C.Structure):
    _fields_ = [("kind", i32), ("reserved", i32), ("col", ColRef), ("c", Literal)]


class ValueExpr(C.Structure):
    _fields_ = [("nfactors", u32), ("reserved", u32), ("factors", Factor * 3)]


class Agg(C.Structure):
    _fields_ = [("func", i32), ("expr", i32)]


class JoinProbe(C.Structure):
    _fields_ = [("join_table", u64), ("probe_key", ColRef)]


class BloomProbe(C.Structure):
    _fields_ = [("bloom", u64), ("expected_generation", u64), ("key", ColRef)]


class SortKey(C.Structure):
    _fields_ = [("is_agg", i32), ("index", i32), ("descending", i32), ("nulls_first", i32)]


class Pipeline(C.Structure):
    _fields_ = [
        ("scan_id", u64),
        ("nbloom", u32), ("bloom", BloomProbe * MAX_BLOOM),
        ("nterms", u32), ("terms", PredTerm * MAX_TERMS),
        ("njoins", u32), ("joins", JoinProbe * MAX_JOINS),
        ("sink", i32),
        ("nkeys", u32), ("keys", ColRef * MAX_KEYS),
        ("nexprs", u32), ("exprs", ValueExpr * MAX_EXPRS),
        ("naggs", u32), ("aggs", Agg * MAX_AGGS),
        ("expected_groups", u64),
        ("build_key", ColRef),
        ("npayload", u32), ("payload", ColRef * MAX_PAYLOAD),
        ("build_bloom", u64),
        ("nsort", u32), ("sort", SortKey * 4),
        ("limit", u64),
        ("build_flags", u32), ("reserved0", u32),
        ("scan_row_set", u64),
    ]


class Value(C.Structure):
    _fields_ = [("kind", i32), ("slen", i32), ("f64", C.c_double), ("lo", i64), ("hi", i64), ("str", u8 * 16)]


class Result(C.Structure):
    _fields_ = [("rows_in", u64), ("rows_bloom", u64), ("rows_filtered", u64), ("rows_out", u64),
                ("ngroups", u64), ("nkeys", u32), ("naggs", u32), ("keys", P(Value)), ("aggs", P(Value)),
                ("join_table", u64), ("bloom_rows", u64), ("kernel_ms", C.c_float), ("kernel_launches", u32),
                ("key_type", i32 * 4), ("agg_type", i32 * 16), ("variant", C.c_char * 24), ("agg_func", i32 * 16),
                ("key_not_null", i32 * 4)]


class RfTarget(C.Structure):
    _fields_ = [("session_epoch", u64), ("scan_id", u64), ("output_column", u32), ("key_type", u32)]


class PoolProbe(C.Structure):
    _fields_ = [("slot_index", i32), ("key_type", u32), ("generation", u64), ("output_column", u32), ("reserved", u32)]


class JoinInfo(C.Structure):
    _fields_ = [("rows", u64), ("capacity", u32), ("row_bytes", u32), ("npayload", u32), ("reserved", u32)]


class GenSpec(C.Structure):
    _fields_ = [("table", i32), ("dense_keys", i32), ("seed", u64), ("first_row", u64), ("rows", u64),
                ("scale_rows", u64)]


_SIGNATURES = {
    # name: (restype, argtypes)
    "pgf_device_count": (i32, [P(i32)]),
    "pgf_ctx_create": (i32, [P(Config), P(vp)]),
    "pgf_ctx_destroy": (None, [vp]),
    "pgf_last_error": (C.c_char_p, [vp]),
    "pgf_ctx_register_host_region": (i32, [vp, vp, C.c_size_t]),
    "pgf_ctx_unregister_host_region": (i32, [vp, vp]),
    "pgf_ctx_synchronize": (i32, [vp]),
    "pgf_ctx_compute_stream": (vp, [vp]),
    "pgf_ctx_last_kernel_ms": (C.c_float, [vp]),
    "pgf_layout_plan_new": (i32, [P(ColumnSpec), u32, u32, u32, P(LayoutPlanC)]),
    "pgf_layout_fixed_row_cap": (i32, [P(ColumnSpec), u32, u32, P(u32)]),
    "pgf_ctx_runtime_filter_metrics": (i32, [vp, P(RuntimeFilterMetrics)]),
    "pgf_ctx_note_pool_exhausted": (i32, [vp]),
    "pgf_block_validate": (i32, [vp, C.c_size_t]),
    "pgf_block_validate_ext": (i32, [vp, C.c_size_t, u32]),
    "pgf_block_import_check": (i32, [u16, u16, vp, C.c_size_t, P(ColumnSpec), u32]),
    "pgf_block_init": (i32, [vp, C.c_size_t, P(LayoutPlanC)]),
    "pgf_block_write_column": (i32, [vp, C.c_size_t, u32, u32, vp, vp]),
    "pgf_block_set_row_count": (i32, [vp, C.c_size_t, u32]),
    "pgf_page_header_encode": (i32, [u16, u16, u32, vp]),
    "pgf_page_header_decode": (i32, [vp, P(u16), P(u16), P(u32)]),
    "pgf_scan_declare": (i32, [vp, u64, P(ColumnSpec), u32, u64]),
    "pgf_scan_push_page": (i32, [vp, u64, vp, u32]),
    "pgf_scan_push_pages": (i32, [vp, u64, vp, u64, u64]),
    "pgf_scan_finish": (i32, [vp, u64]),
    "pgf_scan_get_info": (i32, [vp, u64, P(ScanInfo)]),
    "pgf_scan_reset": (i32, [vp, u64]),
    "pgf_scan_release": (i32, [vp, u64]),
    "pgf_scan_read_pages": (i32, [vp, u64, u64, u64, vp]),
    "pgf_bloom_params_new": (i32, [u64, u64, u64, P(BloomParamsC)]),
    "pgf_bloom_params_for_expected_items": (i32, [u64, C.c_double, u64, P(BloomParamsC)]),
    "pgf_bloom_create": (i32, [vp, P(BloomParamsC), P(u64)]),
    "pgf_bloom_destroy": (i32, [vp, u64]),
    "pgf_bloom_snapshot": (i32, [vp, u64, P(u64), P(i32)]),
    "pgf_bloom_begin_build": (i32, [vp, u64, P(u64)]),
    "pgf_bloom_insert_keys": (i32, [vp, u64, vp, i32, vp, u64, P(u64)]),
    "pgf_bloom_insert_scan": (i32, [vp, u64, u64, u32, P(u64)]),
    "pgf_bloom_publish_ready": (i32, [vp, u64]),
    "pgf_bloom_disable_build": (i32, [vp, u64]),
    "pgf_bloom_retire_ready": (i32, [vp, u64]),
    "pgf_bloom_read_words": (i32, [vp, u64, vp, u64]),
    "pgf_bloom_or_words": (i32, [vp, u64, vp, u64]),
    "pgf_bloom_device_words": (vp, [vp, u64]),
    "pgf_bloom_or_device_words": (i32, [vp, u64, vp, u64, u32]),
    "pgf_bloom_probe_keys": (i32, [vp, u64, u64, vp, i32, vp, u64, vp, P(ProbeStats)]),
    "pgf_bloom_probe_scan": (i32, [vp, u64, u64, u64, u32, vp, P(ProbeStats)]),
    "pgf_pipeline_check": (i32, [vp, P(Pipeline)]),
    "pgf_pipeline_run": (i32, [vp, P(Pipeline), P(P(Result))]),
    "pgf_result_free": (None, [P(Result)]),
    "pgf_join_table_destroy": (i32, [vp, u64]),
    "pgf_pipeline_run_partial": (i32, [vp, P(Pipeline), vp, u64, P(u64), P(P(Result))]),
    "pgf_pipeline_merge_partials": (i32, [vp, P(Pipeline), vp, u64, u32, P(P(Result))]),
    "pgf_partial_state_bytes": (i32, [P(Pipeline), u64, P(u64)]),
    "pgf_shm_pool_layout": (i32, [u32, P(BloomParamsC), P(u64), P(u64)]),
    "pgf_shm_pool_init": (i32, [vp, u64, u32, P(BloomParamsC)]),
    "pgf_shm_pool_attach_check": (i32, [vp, u64, u32, P(BloomParamsC)]),
    "pgf_shm_pool_allocate_build": (i32, [vp, u64, u32, P(BloomParamsC), P(RfTarget), P(i32), P(u64)]),
    "pgf_shm_pool_publish_words": (i32, [vp, u64, u32, P(BloomParamsC), i32, u64, vp, u64]),
    "pgf_shm_pool_disable_build": (i32, [vp, u64, u32, P(BloomParamsC), i32, u64]),
    "pgf_shm_pool_release_owner": (i32, [vp, u64, u32, P(BloomParamsC), i32]),
    "pgf_bloom_publish_to_pool": (i32, [vp, u64, vp, u64, u32, i32, u64]),
    "pgf_shm_pool_lookup_probes": (i32, [vp, u64, u32, P(BloomParamsC), u64, u64, P(PoolProbe), u32, P(u32)]),
    "pgf_shm_pool_probe_decide": (i32, [vp, u64, u32, P(BloomParamsC), i32, u64, i32, i64, P(i32)]),
    "pgf_shm_pool_release_probe": (i32, [vp, u64, u32, P(BloomParamsC), i32]),
    "pgf_result_schema": (i32, [P(Result), P(ColumnSpec), P(u32)]),
    "pgf_result_encode_pages": (i32, [P(Result), u32, u64, vp, u64, P(u64), P(u64)]),
    "pgf_join_table_get_info": (i32, [vp, u64, P(JoinInfo)]),
    "pgf_join_table_export": (i32, [vp, u64, vp, u64, P(u64)]),
    "pgf_join_table_from_fragments": (i32, [vp, u64, vp, u64, P(u64), u32, P(u64)]),
    "pgf_pipeline_run_partial_async": (i32, [vp, P(Pipeline), vp, u64]),
    "pgf_comm_unique_id": (i32, [vp]),
    "pgf_comm_init": (i32, [vp, vp, i32, i32]),
    "pgf_comm_destroy": (i32, [vp]),
    "pgf_comm_info": (i32, [vp, P(i32), P(i32)]),
    "pgf_comm_all_gather": (i32, [vp, vp, vp, u64]),
    "pgf_comm_all_gather_host": (i32, [vp, vp, vp, u64]),
    "pgf_pipeline_run_sharded": (i32, [vp, P(Pipeline), u64, P(P(Result))]),
    "pgf_bloom_or_all_reduce": (i32, [vp, u64]),
    "pgf_join_table_exchange": (i32, [vp, u64, u32, P(u64), P(u64)]),
    "pgf_partition_of_key": (u32, [i64, u32]),
    "pgf_pipeline_merge_partials_bounded": (i32, [vp, P(Pipeline), vp, u64, u32, P(P(Result))]),
    "pgf_gen_scan": (i32, [vp, u64, P(GenSpec)]),
    "pgf_gen_schema": (i32, [i32, P(ColumnSpec), P(u32)]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def build(force: bool = False) -> str:
    """Compile libpgf_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    if force:
        subprocess.check_call(["make", "-C", CSRC_DIR, "clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", CSRC_DIR, "-j8"], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the pg_fusion_b200 hot path)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
