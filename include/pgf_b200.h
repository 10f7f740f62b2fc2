/*
 * pgf_b200.h -- C ABI of libpgf_b200.so: a B200-native (sm_100a) implementation of
 * pg_fusion's worker-side columnar hot path (filter / projection / hash aggregate /
 * int-key hash join / runtime Bloom filter over page-backed Arrow blocks).
 *
 * This is the surface a DataFusion `ExecutionPlan` shim inside pg_fusion's
 * `worker_runtime` binds over FFI.  Every entry point names the
 * reference interface it replaces (file:line relative to the pg_fusion repository).
 *
 * Conventions
 *   - every function returns a pgf_status (0 = PGF_OK); pgf_last_error(ctx) returns a
 *     human readable message for the last failure on that context.  Nothing aborts,
 *     exits or throws across this boundary ("no panics in extension paths",
 *     ai/invariants.md:12-17).  CUDA errors are sticky per context.
 *   - host pointers are borrowed for the duration of the call unless stated otherwise.
 *   - a context is used by one thread at a time, except pgf_scan_push_* and pgf_scan_finish,
 *     which may be called concurrently for different scans (one producer thread per scan,
 *     mirroring worker_runtime/src/transport_scan_source.rs:166-183), also while a pipeline
 *     over other, finished scans runs.
 *   - there is no CPU fallback: without a CUDA device pgf_ctx_create fails.
 */
#ifndef PGF_B200_H
#define PGF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t pgf_status;
enum {
  PGF_OK = 0,
  PGF_ERR_INVALID_ARGUMENT = 1,
  PGF_ERR_CUDA = 2,            /* sticky */
  PGF_ERR_NO_DEVICE = 3,
  PGF_ERR_OUT_OF_MEMORY = 4,
  PGF_ERR_UNKNOWN_HANDLE = 5,
  PGF_ERR_NOT_ELIGIBLE = 6,    /* plan outside the supported grammar: keep the DataFusion node */
  PGF_ERR_STATE = 7,           /* call not valid in the object's current state */
  PGF_ERR_UNSUPPORTED_DATA = 8,/* e.g. out-of-line (> 12 byte) view in a GROUP BY key or join payload column
                                  (predicates compare such values through the page's tail arena) */
  PGF_ERR_COMM = 9,            /* NCCL / communicator error */
  /* BloomParamError / BloomAttachError (runtime_filter/src/bloom.rs:103-137) */
  PGF_ERR_BLOOM_ZERO_BIT_COUNT = 20,
  PGF_ERR_BLOOM_ZERO_HASH_COUNT = 21,
  PGF_ERR_BLOOM_ZERO_EXPECTED_ITEMS = 22,
  PGF_ERR_BLOOM_INVALID_FPR = 23,
  PGF_ERR_BLOOM_TOO_MANY_BITS = 24,
  PGF_ERR_BLOOM_INSUFFICIENT_WORDS = 25,
  /* LifecycleError (runtime_filter/src/shared.rs:58-71) */
  PGF_ERR_LIFECYCLE_GENERATION_EXHAUSTED = 30,
  PGF_ERR_LIFECYCLE_BUSY = 31,
  PGF_ERR_LIFECYCLE_INVALID_TRANSITION = 32,
  /* LayoutError (page/arrow_layout/src/error.rs) */
  PGF_ERR_LAYOUT_INVALID_MAGIC = 101,
  PGF_ERR_LAYOUT_INVALID_VERSION = 102,
  PGF_ERR_LAYOUT_ROW_COUNT_EXCEEDS_MAX_ROWS = 103,
  PGF_ERR_LAYOUT_COLUMN_COUNT_MISMATCH = 104,
  PGF_ERR_LAYOUT_FRONT_BASE_MISMATCH = 105,
  PGF_ERR_LAYOUT_INVALID_HEADER_BOUNDS = 106,
  PGF_ERR_LAYOUT_MISALIGNED_FRONT_REGION = 107,
  PGF_ERR_LAYOUT_BLOCK_SLICE_TOO_SMALL = 108,
  PGF_ERR_LAYOUT_INVALID_TYPE_TAG = 109,
  PGF_ERR_LAYOUT_INCONSISTENT_VIEW_FLAG = 110,
  PGF_ERR_LAYOUT_COLUMN_DESC_MISMATCH = 111,
  PGF_ERR_LAYOUT_POOL_BASE_MISMATCH = 112,
  PGF_ERR_LAYOUT_DOES_NOT_FIT = 113,
  PGF_ERR_LAYOUT_SIZE_OVERFLOW = 114,
  PGF_ERR_LAYOUT_TOO_MANY_COLUMNS = 115,
  PGF_ERR_LAYOUT_NEGATIVE_VIEW_LENGTH = 117,
  PGF_ERR_LAYOUT_INVALID_VIEW_BUFFER_INDEX = 118,
  PGF_ERR_LAYOUT_NEGATIVE_VIEW_OFFSET = 119,
  PGF_ERR_LAYOUT_VIEW_OFFSET_OUT_OF_BOUNDS = 120,
  PGF_ERR_LAYOUT_VIEW_FULL = 122,
  /* ImportError (page/import/src/error.rs) */
  PGF_ERR_IMPORT_WRONG_KIND = 201,
  PGF_ERR_IMPORT_UNSUPPORTED_FLAGS = 202,
  PGF_ERR_IMPORT_SCHEMA_COLUMN_COUNT_MISMATCH = 203,
  PGF_ERR_IMPORT_SCHEMA_TYPE_MISMATCH = 204,
  PGF_ERR_IMPORT_SCHEMA_NULLABILITY_MISMATCH = 205,
  PGF_ERR_IMPORT_INVALID_NULL_COUNT = 206,
  PGF_ERR_IMPORT_NULL_BITMAP_COUNT_MISMATCH = 207,
  PGF_ERR_IMPORT_VIEW_OFFSET_BEFORE_ALLOCATED_TAIL = 208,
  PGF_ERR_IMPORT_PAGE_HEADER_INVALID = 209,
  PGF_ERR_IMPORT_ARROW_INVALID_VIEW = 210
};

/* On-page type tags: page/arrow_layout/src/types.rs:93-112.  PGF_T_DECIMAL128 is an
 * extension beyond reference v1 pages (16-byte little-endian two's complement slot). */
enum {
  PGF_T_BOOLEAN = 1, PGF_T_INT16 = 2, PGF_T_INT32 = 3, PGF_T_INT64 = 4, PGF_T_FLOAT32 = 5,
  PGF_T_FLOAT64 = 6, PGF_T_UUID = 7, PGF_T_UTF8VIEW = 8, PGF_T_BINARYVIEW = 9,
  PGF_T_DECIMAL128 = 10
};

#define PGF_PAGE_HEADER_LEN 20u          /* page/transfer/src/page.rs:11 */
#define PGF_ARROW_LAYOUT_BATCH_KIND 0x4152u /* page/import/src/lib.rs:43 */
#define PGF_MAX_COLS 16u                 /* columns per scan this library stages */

typedef struct pgf_ctx pgf_ctx;

/* Replaces the relevant part of HostConfig (pg/extension/src/guc.rs:47-75): page_size is
 * pg_fusion.page_size (guc.rs:31-32; the block payload is page_size - 20). */
typedef struct {
  int32_t device;          /* CUDA device ordinal */
  uint32_t page_size;      /* 0 => 65536 */
  uint32_t staging_pages;  /* pinned staging ring for unregistered host pages; 0 => 512 */
  uint32_t flags;          /* PGF_CFG_* */
} pgf_config;
/* A runtime filter is an optimisation only: rejecting early what a later join rejects anyway
 * (runtime_filter/src/shared.rs:350-374 allows PassUnfiltered at any time).  By default the fused
 * pipelines therefore drop a Bloom probe (a) when the same pipeline probes a join table on the same
 * key column -- the table's tag directory already rejects misses at the same cost -- and (b) when
 * the filter is saturated (expected pass rate of absent keys fill^k > 0.9).  This flag keeps every
 * probe (tests of the fused probe itself, metrics parity with the backend-side probe counters). */
#define PGF_CFG_KEEP_REDUNDANT_BLOOM_PROBES 1u

/* ----------------------------------------------------------------- context */
pgf_status pgf_device_count(int32_t *count_out);
pgf_status pgf_ctx_create(const pgf_config *config, pgf_ctx **ctx_out);
void pgf_ctx_destroy(pgf_ctx *ctx);
const char *pgf_last_error(const pgf_ctx *ctx);
/* Pin a host region (e.g. the shared-memory page pool, page/pool) so pages inside it are
 * DMA-copied without the staging memcpy. */
pgf_status pgf_ctx_register_host_region(pgf_ctx *ctx, void *base, size_t len);
pgf_status pgf_ctx_unregister_host_region(pgf_ctx *ctx, void *base);
/* Block until all work queued on the context's streams has completed. */
pgf_status pgf_ctx_synchronize(pgf_ctx *ctx);
/* Device time (CUDA events on the compute stream) of the kernels of the last Bloom build /
 * probe or pipeline call on this context, in milliseconds. */
float pgf_ctx_last_kernel_ms(const pgf_ctx *ctx);
/* The compute stream (cudaStream_t) so a harness can bracket it with CUDA events. */
void *pgf_ctx_compute_stream(pgf_ctx *ctx);

/* ---------------------------------------------------- page layout (host side) */
typedef struct { uint16_t type_tag; uint16_t nullable; } pgf_column_spec; /* types.rs:218-225 */
typedef struct {
  uint16_t type_tag; uint16_t flags;
  uint32_t validity_off, values_off, validity_len, values_len;
} pgf_column_layout;                                                        /* types.rs:250-264 */
typedef struct {
  uint32_t block_size, max_rows, front_base, pool_base, ncols;
  pgf_column_layout cols[64];
} pgf_layout_plan;                                                          /* plan.rs:22-29 */

/* LayoutPlan::new, page/arrow_layout/src/plan.rs:33-93 */
pgf_status pgf_layout_plan_new(const pgf_column_spec *specs, uint32_t ncols, uint32_t max_rows,
                               uint32_t block_size, pgf_layout_plan *plan_out);
/* compute_fixed_row_cap, page/row_estimator/src/lib.rs:353-371 */
pgf_status pgf_layout_fixed_row_cap(const pgf_column_spec *specs, uint32_t ncols,
                                    uint32_t block_size, uint32_t *cap_out);
/* BlockRef::open, page/arrow_layout/src/access.rs:36-42.  Strict reference v1: type tags 1..9 only
 * (TypeTag::from_raw, types.rs:93-112); a block carrying the Decimal128 extension tag is
 * PGF_ERR_LAYOUT_INVALID_TYPE_TAG here, exactly as in the reference. */
pgf_status pgf_block_validate(const uint8_t *block, size_t len);
/* The same check with this library's format extensions switched on explicitly. */
#define PGF_LAYOUT_EXT_DECIMAL128 1u
pgf_status pgf_block_validate_ext(const uint8_t *block, size_t len, uint32_t extensions);
/* ArrowPageDecoder::import_owned checks, page/import/src/lib.rs:117-206 (all of them,
 * host side; the scan ingest path below splits them between host and device).  The Decimal128
 * extension tag passes only under a schema that names it. */
pgf_status pgf_block_import_check(uint16_t kind, uint16_t flags, const uint8_t *block, size_t len,
                                  const pgf_column_spec *schema, uint32_t ncols);
/* init_block, page/arrow_layout/src/access.rs:640-654; then bulk column writes in the
 * style of page/batch_encoder/src/encoder.rs:49-298 (fixed-width values, inline views).
 * values: nrows * width bytes (views: nrows 16-byte ByteView slots); validity: LSB-first
 * bitmap or NULL (all valid). */
pgf_status pgf_block_init(uint8_t *block, size_t len, const pgf_layout_plan *plan);
pgf_status pgf_block_write_column(uint8_t *block, size_t len, uint32_t col, uint32_t nrows,
                                  const void *values, const uint8_t *validity);
pgf_status pgf_block_set_row_count(uint8_t *block, size_t len, uint32_t nrows);
/* transfer page header, page/transfer/src/page.rs:20-64 */
pgf_status pgf_page_header_encode(uint16_t kind, uint16_t flags, uint32_t payload_len,
                                  uint8_t out[20]);
pgf_status pgf_page_header_decode(const uint8_t in[20], uint16_t *kind, uint16_t *flags,
                                  uint32_t *payload_len);

/* --------------------------------------------------------------------- scans
 * Replaces WorkerPgScanExec + ArrowPageDecoder + PageMaterializeExec
 * (worker_runtime/src/scan_exec.rs:139-268; page/import/src/lib.rs:117-206;
 * pg/scan_node/src/page_materialize.rs:107-207): pages are validated, copied to HBM
 * (which also satisfies the "deep copy before retaining operators" rule) and become the
 * input of pipelines. */
pgf_status pgf_scan_declare(pgf_ctx *ctx, uint64_t scan_id, const pgf_column_spec *schema,
                            uint32_t ncols, uint64_t expected_pages);
/* One whole transfer page: 20-byte header + arrow_layout block (len <= page_size). */
pgf_status pgf_scan_push_page(pgf_ctx *ctx, uint64_t scan_id, const uint8_t *page, uint32_t len);
/* npages pages `stride` bytes apart.  If the memory is registered/pinned the copy is
 * asynchronous and the pages must stay valid until pgf_scan_finish returns. */
pgf_status pgf_scan_push_pages(pgf_ctx *ctx, uint64_t scan_id, const uint8_t *pages,
                               uint64_t npages, uint64_t stride);
/* End of stream: waits for the copies and runs the row-level import checks (null bitmap
 * popcounts, view validation; page/import/src/lib.rs:237-293,424-452) on the device. */
pgf_status pgf_scan_finish(pgf_ctx *ctx, uint64_t scan_id);
typedef struct { uint64_t pages, rows, bytes; uint32_t ncols, finished; } pgf_scan_info;
pgf_status pgf_scan_get_info(pgf_ctx *ctx, uint64_t scan_id, pgf_scan_info *out);
/* Drop the scan's pages but keep its declaration (for re-ingest of the same stream). */
pgf_status pgf_scan_reset(pgf_ctx *ctx, uint64_t scan_id);
pgf_status pgf_scan_release(pgf_ctx *ctx, uint64_t scan_id);
/* Copy the device-resident pages back to the host (tests, debugging). */
pgf_status pgf_scan_read_pages(pgf_ctx *ctx, uint64_t scan_id, uint64_t first_page,
                               uint64_t npages, uint8_t *out);

/* ------------------------------------------------------------- Bloom filter
 * Replaces runtime_filter (BloomParams, AtomicBloomRef, RuntimeFilterSlot lifecycle):
 * runtime_filter/src/bloom.rs:17-100,159-256; runtime_filter/src/shared.rs:132-416. */
typedef struct { uint64_t bit_count, word_count, hash_count, seed; } pgf_bloom_params;
pgf_status pgf_bloom_params_new(uint64_t bit_count, uint64_t hash_count, uint64_t seed,
                                pgf_bloom_params *out);                 /* bloom.rs:29-48 */
pgf_status pgf_bloom_params_for_expected_items(uint64_t expected_items, double fpr,
                                               uint64_t seed, pgf_bloom_params *out); /* :52-79 */

enum { PGF_RF_FREE = 0, PGF_RF_BUILDING = 1, PGF_RF_READY = 2, PGF_RF_DISABLED = 3 };
enum { PGF_PASS_UNFILTERED = 0, PGF_MAYBE_PRESENT = 1, PGF_DEFINITELY_ABSENT = 2 };

/* A filter slot in HBM with the reference's lifecycle word (generation << 2 | state). */
pgf_status pgf_bloom_create(pgf_ctx *ctx, const pgf_bloom_params *params, uint64_t *bloom_out);
pgf_status pgf_bloom_destroy(pgf_ctx *ctx, uint64_t bloom);
pgf_status pgf_bloom_snapshot(pgf_ctx *ctx, uint64_t bloom, uint64_t *generation, int32_t *state);
/* try_acquire_builder (shared.rs:159-198): Free|Disabled -> Building, clears the bits. */
pgf_status pgf_bloom_begin_build(pgf_ctx *ctx, uint64_t bloom, uint64_t *generation_out);
/* RuntimeFilterBuildStream::insert_batch / insert_ints
 * (worker_runtime/src/runtime_filter_plan.rs:227-274,345-363): insert every non-null key
 * of a host key array (key_width 2/4/8, sign-extended) or of one column of a scan. */
pgf_status pgf_bloom_insert_keys(pgf_ctx *ctx, uint64_t bloom, const void *keys, int32_t key_width,
                                 const uint8_t *validity, uint64_t n, uint64_t *rows_inserted);
pgf_status pgf_bloom_insert_scan(pgf_ctx *ctx, uint64_t bloom, uint64_t scan_id, uint32_t col,
                                 uint64_t *rows_inserted);
pgf_status pgf_bloom_publish_ready(pgf_ctx *ctx, uint64_t bloom);  /* shared.rs:201-208 */
pgf_status pgf_bloom_disable_build(pgf_ctx *ctx, uint64_t bloom);  /* shared.rs:211-213 */
pgf_status pgf_bloom_retire_ready(pgf_ctx *ctx, uint64_t bloom);   /* shared.rs:244-260 */
/* Word array exchange: read the bits (bit layout = bloom.rs:243-247: word = bit / 64,
 * mask = 1 << (bit % 64)), OR another array into them (multi-GPU merge / shm interop). */
pgf_status pgf_bloom_read_words(pgf_ctx *ctx, uint64_t bloom, uint64_t *words_out, uint64_t nwords); /* host or device buffer */
pgf_status pgf_bloom_or_words(pgf_ctx *ctx, uint64_t bloom, const uint64_t *words, uint64_t nwords);
/* device pointer to the words (for NCCL all-gather by the harness) */
void *pgf_bloom_device_words(pgf_ctx *ctx, uint64_t bloom);
pgf_status pgf_bloom_or_device_words(pgf_ctx *ctx, uint64_t bloom, const void *dev_words,
                                     uint64_t nwords, uint32_t narrays);
/* Probe (runtime_filter_rejects_slot, pg/backend_service/src/source.rs:496-532;
 * decision_for_hash / decision_for_null, shared.rs:350-374): decisions[i] is a
 * PGF_PASS_UNFILTERED / PGF_MAYBE_PRESENT / PGF_DEFINITELY_ABSENT byte per row. */
typedef struct { uint64_t probe_rows, rejected_rows, pass_unfiltered; } pgf_probe_stats;

/* The runtime-filter counters of the reference's metrics registry (runtime_metrics/src/lib.rs:125-131), accumulated
 * per context: where worker_runtime and backend_service bump a MetricId, the library bumps the field of the same name.
 *   allocated        pgf_bloom_begin_build succeeded (runtime_filter_plan.rs:92)
 *   ready            pgf_bloom_publish_ready succeeded (runtime_filter_plan.rs:283)
 *   pool_exhausted   the planner hook found no free slot and kept the plain join (runtime_filter_plan.rs:89): the hook
 *                    reports it with pgf_ctx_note_pool_exhausted
 *   build_rows       keys inserted (runtime_filter_plan.rs:272): pgf_bloom_insert_*, pgf_pipeline.build_bloom
 *   probe_rows / probe_rows_rejected / probe_pass_unfiltered
 *                    rows tested against a filter, rows it rejected, and rows that met a filter the probe could not
 *                    consult -- not Ready, another generation, or dropped by the lowering as redundant
 *                    (backend_service/src/source.rs:476-488): pgf_bloom_probe_*, fused probes of pgf_pipeline_run */
typedef struct {
  uint64_t allocated_total, ready_total, pool_exhausted_total, build_rows_total;
  uint64_t probe_rows_total, probe_rows_rejected_total, probe_pass_unfiltered_total;
} pgf_runtime_filter_metrics;
pgf_status pgf_ctx_runtime_filter_metrics(pgf_ctx *ctx, pgf_runtime_filter_metrics *out);
pgf_status pgf_ctx_note_pool_exhausted(pgf_ctx *ctx);
pgf_status pgf_bloom_probe_keys(pgf_ctx *ctx, uint64_t bloom, uint64_t expected_generation,
                                const void *keys, int32_t key_width, const uint8_t *validity,
                                uint64_t n, uint8_t *decisions_out, pgf_probe_stats *stats);
pgf_status pgf_bloom_probe_scan(pgf_ctx *ctx, uint64_t bloom, uint64_t expected_generation,
                                uint64_t scan_id, uint32_t col, uint8_t *decisions_out,
                                pgf_probe_stats *stats);

/* ---- shared-memory runtime-filter pool (runtime_filter/src/pool.rs)
 * The worker side of the reference's pool protocol, so a filter built on the GPU is probed by
 * unchanged PostgreSQL backends before tuples are encoded (pg/backend_service/src/source.rs:
 * 121-149,496-532).  `base` is the mapped pool region; (slot_count, params) is the pool
 * configuration (RuntimeFilterPoolConfig).  allocate_build = RuntimeFilterPool::allocate_build
 * (pool.rs:378-430; *slot_index_out = -1 when the pool is exhausted, a soft miss),
 * publish = insert + RuntimeFilterBuildHandle::publish_ready, release_owner = handle drop. */
typedef struct { uint64_t session_epoch, scan_id; uint32_t output_column, key_type; /* 1=i16 2=i32 3=i64 */ } pgf_rf_target;
pgf_status pgf_shm_pool_layout(uint32_t slot_count, const pgf_bloom_params *params, uint64_t *size_out, uint64_t *align_out);
pgf_status pgf_shm_pool_init(void *base, uint64_t len, uint32_t slot_count, const pgf_bloom_params *params);
pgf_status pgf_shm_pool_attach_check(void *base, uint64_t len, uint32_t slot_count, const pgf_bloom_params *params);
pgf_status pgf_shm_pool_allocate_build(void *base, uint64_t len, uint32_t slot_count, const pgf_bloom_params *params,
                                       const pgf_rf_target *target, int32_t *slot_index_out, uint64_t *generation_out);
pgf_status pgf_shm_pool_publish_words(void *base, uint64_t len, uint32_t slot_count, const pgf_bloom_params *params,
                                      int32_t slot_index, uint64_t generation, const uint64_t *words, uint64_t nwords);
pgf_status pgf_shm_pool_disable_build(void *base, uint64_t len, uint32_t slot_count, const pgf_bloom_params *params,
                                      int32_t slot_index, uint64_t generation);
pgf_status pgf_shm_pool_release_owner(void *base, uint64_t len, uint32_t slot_count, const pgf_bloom_params *params,
                                      int32_t slot_index);
/* Probe side of the same protocol (lookup_probes / decision_for_hash / handle drop, pool.rs:432-476 and
 * shared.rs:350-374).  Backends run the reference's own code for this; it is exported so the pool
 * implementation can be checked against the reference's pool tests. */
typedef struct { int32_t slot_index; uint32_t key_type; uint64_t generation; uint32_t output_column, reserved; } pgf_pool_probe;
pgf_status pgf_shm_pool_lookup_probes(void *base, uint64_t len, uint32_t slot_count, const pgf_bloom_params *params,
                                      uint64_t session_epoch, uint64_t scan_id, pgf_pool_probe *out,
                                      uint32_t max_probes, uint32_t *nprobes_out);
pgf_status pgf_shm_pool_probe_decide(void *base, uint64_t len, uint32_t slot_count, const pgf_bloom_params *params,
                                     int32_t slot_index, uint64_t generation, int32_t key_is_null, int64_t key,
                                     int32_t *decision_out);
pgf_status pgf_shm_pool_release_probe(void *base, uint64_t len, uint32_t slot_count, const pgf_bloom_params *params,
                                      int32_t slot_index);
/* Copy the words of a GPU-built filter (any lifecycle state) into the pool slot and publish it. */
pgf_status pgf_bloom_publish_to_pool(pgf_ctx *ctx, uint64_t bloom, void *base, uint64_t len, uint32_t slot_count,
                                     int32_t slot_index, uint64_t generation);

/* ------------------------------------------------------------------ pipelines
 * Replaces the operator chain DataFusion plans over a scan
 * (worker_runtime/src/runtime.rs:667-698): CoalesceBatchesExec/FilterExec ->
 * [HashJoinExec probe side]* -> AggregateExec(mode=Single) | HashJoinExec build side
 * (+ RuntimeFilterBuildExec).  One pipeline = one fused kernel over the scan's pages.
 * Plans outside this grammar return PGF_ERR_NOT_ELIGIBLE (the shim keeps the DataFusion
 * node, exactly like install_runtime_filters skips ineligible joins,
 * worker_runtime/src/runtime_filter_plan.rs:50-111). */
enum { PGF_CMP_LT = 0, PGF_CMP_LE = 1, PGF_CMP_GT = 2, PGF_CMP_GE = 3, PGF_CMP_EQ = 4, PGF_CMP_NE = 5 };

typedef struct {
  int32_t type_tag;   /* PGF_T_FLOAT64 / PGF_T_INT64 / PGF_T_UTF8VIEW / PGF_T_DECIMAL128 / PGF_T_BOOLEAN (i64 = 0 / 1) ... */
  int32_t slen;       /* string literal length (<= 12) */
  double f64;
  int64_t i64;        /* integer literal, or low 64 bits of a Decimal128 literal */
  int64_t hi;         /* high 64 bits of a Decimal128 literal */
  uint8_t str[16];
} pgf_literal;

/* Reference to a value: column `col` of source 0 (the pipeline's scan), or payload slot
 * `col` of the build side of join `source - 1`. */
typedef struct { int32_t source; int32_t col; } pgf_colref;

/* One conjunct of the FilterExec predicate: <column> <cmp> <literal>.  The predicate is
 * the AND of all terms; a row passes iff every term is TRUE (NULL => dropped).  Columns: the
 * integer and float types, Decimal128, Utf8View / BinaryView (values of any length -- out-of-line
 * ones are read from the page's tail arena, page/arrow_layout/src/raw.rs:98-110 -- against
 * literals of at most 12 bytes) and Boolean (bit-packed values, types.rs:139-147; literal
 * PGF_T_BOOLEAN: `WHERE flag` is flag = true). */
typedef struct { pgf_colref col; int32_t cmp; int32_t reserved; pgf_literal lit; } pgf_pred_term;

/* Projection / aggregate argument: a product of up to 3 factors, each `x`, `(c - x)` or
 * `(c + x)` -- covers l_extendedprice, price*discount, price*(1-discount),
 * price*(1-discount)*(1+tax) (benches/tpch/queries/q01.sql,q03.sql,q06.sql).  Evaluated per
 * row exactly as written: IEEE Float64 without FMA contraction, or wrapping i64 / i128. */
enum { PGF_FACTOR_COL = 0, PGF_FACTOR_CONST_MINUS_COL = 1, PGF_FACTOR_CONST_PLUS_COL = 2 };
typedef struct { int32_t kind; int32_t reserved; pgf_colref col; pgf_literal c; } pgf_factor;
typedef struct { uint32_t nfactors; uint32_t reserved; pgf_factor factors[3]; } pgf_value_expr;

enum { PGF_AGG_SUM = 1, PGF_AGG_AVG = 2, PGF_AGG_COUNT_STAR = 3, PGF_AGG_COUNT = 4 };
typedef struct { int32_t func; int32_t expr; /* index into exprs; ignored for COUNT(*) */ } pgf_agg;

typedef struct { uint64_t join_table; pgf_colref probe_key; } pgf_join_probe;
typedef struct { uint64_t bloom; uint64_t expected_generation; pgf_colref key; } pgf_bloom_probe;

enum {
  PGF_SINK_AGGREGATE = 1,  /* AggregateExec(mode=Single): nkeys == 0 => one output row */
  PGF_SINK_JOIN_BUILD = 2, /* HashJoinExec build side (CollectLeft) [+ Bloom build] */
  PGF_SINK_COUNT = 3       /* count the rows that reach the sink (testing / EXPLAIN ANALYZE) */
};

#define PGF_MAX_TERMS 8u
#define PGF_MAX_JOINS 2u
#define PGF_MAX_BLOOM_PROBES 2u
#define PGF_MAX_KEYS 4u
#define PGF_MAX_EXPRS 8u
#define PGF_MAX_AGGS 16u
#define PGF_MAX_PAYLOAD 4u
#define PGF_MAX_SORT 4u
#define PGF_TOPK_DEVICE_MAX 64u

/* One ORDER BY term: output column (group key `index` or aggregate `index`), direction, NULL order. */
typedef struct { int32_t is_agg; int32_t index; int32_t descending; int32_t nulls_first; } pgf_sort_key;

typedef struct {
  uint64_t scan_id;
  uint32_t nbloom;  pgf_bloom_probe bloom[PGF_MAX_BLOOM_PROBES];
  uint32_t nterms;  pgf_pred_term terms[PGF_MAX_TERMS];
  uint32_t njoins;  pgf_join_probe joins[PGF_MAX_JOINS];
  int32_t sink;
  /* PGF_SINK_AGGREGATE */
  uint32_t nkeys;   pgf_colref keys[PGF_MAX_KEYS];
  uint32_t nexprs;  pgf_value_expr exprs[PGF_MAX_EXPRS];
  uint32_t naggs;   pgf_agg aggs[PGF_MAX_AGGS];
  uint64_t expected_groups;       /* sizing hint: groups of an aggregate sink / build rows of a build sink; 0 = unknown */
  /* PGF_SINK_JOIN_BUILD: key must be Int16/Int32/Int64; payload columns are carried in the
   * table and addressed by later pipelines as pgf_colref{source = join index + 1, col = i} */
  pgf_colref build_key;
  uint32_t npayload; pgf_colref payload[PGF_MAX_PAYLOAD];
  uint64_t build_bloom;           /* 0 = none; else a filter in Building state to populate */
  /* PGF_SINK_AGGREGATE: SortExec / TopK above the aggregate (ORDER BY ... [LIMIT n], e.g.
   * benches/tpch/queries/q03.sql "ORDER BY revenue DESC, o_orderdate LIMIT 10").  Output rows are
   * returned in this order; with limit <= PGF_TOPK_DEVICE_MAX the selection runs on the device
   * and only `limit` rows leave the GPU.  DataFusion defaults: ASC NULLS LAST, DESC NULLS FIRST. */
  uint32_t nsort;  pgf_sort_key sort[PGF_MAX_SORT];
  uint64_t limit;                 /* 0 = no limit */
  /* PGF_SINK_JOIN_BUILD: PGF_BUILD_ROWS_ONLY keeps the build rows as a dense ROW SET and builds no hash table
   * (result.join_table is then a row-set handle: input of pgf_join_table_exchange or of a row-set scan). */
  uint32_t build_flags;
  uint32_t reserved0;
  /* != 0: the pipeline scans this row set instead of scan_id (the probe side of a hash-partitioned join
   * after the exchange).  Column 0 of source 0 is the row's key, column i + 1 its payload i; no Bloom probes
   * and no predicate terms. */
  uint64_t scan_row_set;
} pgf_pipeline;
#define PGF_BUILD_ROWS_ONLY 1u

/* Result values */
enum { PGF_V_NULL = 0, PGF_V_F64 = 1, PGF_V_I64 = 2, PGF_V_I128 = 3, PGF_V_STR = 4 };
typedef struct {
  int32_t kind; int32_t slen;
  double f64;
  int64_t lo, hi;      /* i64 in lo; i128 in (hi:lo) */
  uint8_t str[16];
} pgf_value;

typedef struct {
  uint64_t rows_in;        /* rows scanned */
  uint64_t rows_bloom;     /* rows_in minus the rows a runtime filter rejected (RuntimeFilterProbeRowsRejectedTotal);
                            * pipelines with joins / build sinks probe only rows the predicate kept */
  uint64_t rows_filtered;  /* rows surviving the runtime filters and the predicate */
  uint64_t rows_out;       /* rows reaching the sink (after joins) */
  uint64_t ngroups;        /* PGF_SINK_AGGREGATE: output rows */
  uint32_t nkeys, naggs;
  pgf_value *keys;         /* ngroups * nkeys  (group order is unspecified) */
  pgf_value *aggs;         /* ngroups * naggs */
  uint64_t join_table;     /* PGF_SINK_JOIN_BUILD: handle of the built table */
  uint64_t bloom_rows;     /* keys inserted into build_bloom (RuntimeFilterBuildRowsTotal) */
  float kernel_ms;         /* device time of the fused kernel(s), CUDA events */
  uint32_t kernel_launches;
  /* Transport types of the output columns (PGF_T_*), keys first, then aggregates: group keys
   * keep their page type; SUM/AVG(Float64) and AVG(int) -> Float64, SUM(int) and COUNT -> Int64,
   * SUM/AVG(Decimal128) -> Decimal128 (same mapping as DataFusion's aggregate return types
   * after worker_runtime/src/result_pages.rs:201-249 normalize_result_transport_schema). */
  int32_t key_type[PGF_MAX_KEYS];
  int32_t agg_type[PGF_MAX_AGGS];
  char variant[24];        /* which instantiation of the fused kernel ran: a registered shape or "generic" (EXPLAIN ANALYZE aid) */
  /* PGF_AGG_* of every aggregate: decides the nullability of its result column (COUNT is NOT NULL,
   * SUM / AVG are nullable whatever the data holds).  0 in a hand-built result = unknown. */
  int32_t agg_func[PGF_MAX_AGGS];
  /* 1: group key k comes from a NOT NULL column (of the scan, or of a join build side), so its output
   * field is non-nullable -- exactly what DataFusion's AggregateExec derives from the input schema and
   * what ArrowPageDecoder::validate_schema compares (page/import/src/lib.rs:208-235; a nullability
   * mismatch is SchemaNullabilityMismatch).  0: nullable, or unknown in a hand-built result. */
  int32_t key_not_null[PGF_MAX_KEYS];
} pgf_result;

pgf_status pgf_pipeline_check(pgf_ctx *ctx, const pgf_pipeline *plan); /* eligibility only */
pgf_status pgf_pipeline_run(pgf_ctx *ctx, const pgf_pipeline *plan, pgf_result **result_out);
void pgf_result_free(pgf_result *result);

/* Result pages.  Replaces ResultPageProducer::encode_pending_page + BatchPageEncoder
 * (worker_runtime/src/result_pages.rs:150-196; page/batch_encoder/src/encoder.rs:49-298) for
 * the aggregate output of a pipeline: rows are encoded into transfer pages (20-byte header, kind
 * 0x4152, payload = one arrow_layout block of page_size - 20 bytes with max_rows = the fixed row
 * cap of the schema) that slot_import / ArrowPageDecoder consume unchanged.  A group key column is
 * nullable iff its source column is (key_not_null); SUM / AVG are nullable, COUNT is not -- the output
 * schema DataFusion derives for AggregateExec.  Strings are inline views (<= 12 bytes).
 * `first_row` / `max_pages` allow page-at-a-time production like next_step(); *rows_done is the
 * number of rows encoded by this call. */
pgf_status pgf_result_schema(const pgf_result *result, pgf_column_spec *schema_out, uint32_t *ncols_out);
pgf_status pgf_result_encode_pages(const pgf_result *result, uint32_t page_size, uint64_t first_row,
                                   uint8_t *pages_out, uint64_t max_pages, uint64_t *npages_out,
                                   uint64_t *rows_done);
pgf_status pgf_join_table_destroy(pgf_ctx *ctx, uint64_t join_table);

/* Broadcast join across GPUs (HashJoinExec CollectLeft with the build side sharded by page;
 * SURVEY.md 8e): every rank builds a table from its own pages, exports its rows as
 * self-contained slot records (row_bytes each: key, NULL flags, payload), the host all-gathers
 * the fragments (NCCL) and every rank rebuilds the full table from them.  `like_table`
 * supplies the key / payload schema. */
typedef struct { uint64_t rows; uint32_t capacity, row_bytes, npayload, reserved; } pgf_join_info;
pgf_status pgf_join_table_get_info(pgf_ctx *ctx, uint64_t join_table, pgf_join_info *out);
pgf_status pgf_join_table_export(pgf_ctx *ctx, uint64_t join_table, void *dev_rows_out,
                                 uint64_t capacity_rows, uint64_t *rows_out);
pgf_status pgf_join_table_from_fragments(pgf_ctx *ctx, uint64_t like_table, const void *dev_rows,
                                         uint64_t stride_bytes, const uint64_t *counts,
                                         uint32_t nfragments, uint64_t *table_out);

/* Partial / Final aggregation across GPUs (AggregateExec Partial -> FinalPartitioned):
 * run the pipeline but leave the per-group partial states in a device buffer; gather the
 * buffers of all ranks (NCCL all-gather by the host) and merge them in rank order into a
 * final result.  pgf_partial_state_bytes gives an upper bound of the buffer size for
 * `max_groups` groups. */
pgf_status pgf_partial_state_bytes(const pgf_pipeline *plan, uint64_t max_groups, uint64_t *bytes_out);
pgf_status pgf_pipeline_run_partial(pgf_ctx *ctx, const pgf_pipeline *plan, void *dev_state_out,
                                    uint64_t state_capacity_bytes, uint64_t *state_bytes_out,
                                    pgf_result **stats_out);
pgf_status pgf_pipeline_merge_partials(pgf_ctx *ctx, const pgf_pipeline *plan,
                                       const void *dev_states, uint64_t state_stride_bytes,
                                       uint32_t nstates, pgf_result **result_out);

/* One-synchronisation form of the multi-GPU step.  pgf_pipeline_run_partial_async enqueues the
 * fused kernel and the extraction of the partial state on the compute stream and returns
 * without synchronising (errors of the run surface in the merge).  After the all-gather has
 * been enqueued on the same stream, pgf_pipeline_merge_partials_bounded sizes the final table
 * from what the strides can hold instead of reading the counts back, merges in rank order
 * and synchronises once for the result. */
pgf_status pgf_pipeline_run_partial_async(pgf_ctx *ctx, const pgf_pipeline *plan, void *dev_state_out,
                                          uint64_t state_capacity_bytes);
pgf_status pgf_pipeline_merge_partials_bounded(pgf_ctx *ctx, const pgf_pipeline *plan,
                                               const void *dev_states, uint64_t state_stride_bytes,
                                               uint32_t nstates, pgf_result **result_out);

/* ------------------------------------------------------------ multi-GPU (one process per GPU)
 * The reference runs every operator in one partition (worker_runtime/src/runtime.rs:748-758) and moves pages
 * between processes through shared memory only; across the GPUs of one box this library shards every scan by
 * pages (the analogue of the reference's CTID-range scan producers, ai/architecture.md:119-133) and runs the
 * exchange steps itself, with NCCL over NVLink / NVSwitch (libnccl.so.2 is loaded on first use; a single-GPU host
 * never needs it).  One context per process and GPU; rank 0 creates the id and hands it to the other ranks out
 * of band (shared memory in the worker, a file or torch.distributed in the tests). */
#define PGF_COMM_ID_BYTES 128
pgf_status pgf_comm_unique_id(uint8_t id_out[PGF_COMM_ID_BYTES]);
pgf_status pgf_comm_init(pgf_ctx *ctx, const uint8_t id[PGF_COMM_ID_BYTES], int32_t rank, int32_t world);
pgf_status pgf_comm_destroy(pgf_ctx *ctx);
pgf_status pgf_comm_info(pgf_ctx *ctx, int32_t *rank_out, int32_t *world_out);
/* all-gather of `bytes` device bytes per rank on the context's compute stream (stream ordered, not synchronised) */
pgf_status pgf_comm_all_gather(pgf_ctx *ctx, const void *dev_send, void *dev_recv, uint64_t bytes);
/* the same for small host buffers (row counts, the top-k rows of every rank): staged through the device, synchronous */
pgf_status pgf_comm_all_gather_host(pgf_ctx *ctx, const void *host_send, void *host_recv, uint64_t bytes);
/* AggregateExec Partial -> exchange -> Final in one call: the fused kernel over this rank's pages, extraction of
 * the partial state (sized for max_groups groups: 72 bytes for Q6), all-gather of the states, fixed-order merge
 * (rank order: every rank gets the same bits; exact for Int64 / Decimal128) and one synchronisation for the result. */
pgf_status pgf_pipeline_run_sharded(pgf_ctx *ctx, const pgf_pipeline *plan, uint64_t max_groups, pgf_result **result_out);
/* Bloom OR-merge: all-gather of the word arrays of every rank's filter (Building state) + bitwise OR;
 * bit exact whatever the order.  Stream ordered. */
pgf_status pgf_bloom_or_all_reduce(pgf_ctx *ctx, uint64_t bloom);
/* Join exchange.  Input: a join table or a row set built from this rank's pages.
 * PGF_XCHG_BROADCAST: every rank receives every row (all-gather) -- small build sides.
 * PGF_XCHG_PARTITION: row r goes to rank hash(key) % world (all-to-all, grouped ncclSend / ncclRecv) -- large build
 *   sides, and the probe side of a partitioned join.  Rows with equal keys meet on one rank, so a GROUP BY that
 *   contains the join key needs no further merge.
 * PGF_XCHG_ROWS_ONLY: the output is a row set (no hash table): the input of a row-set scan.
 * The input handle stays valid.  nvlink_bytes_out (optional): bytes this rank sent over NVLink. */
enum { PGF_XCHG_BROADCAST = 0, PGF_XCHG_PARTITION = 1, PGF_XCHG_ROWS_ONLY = 4 };
/* Tests on one GPU: a context WITHOUT a communicator plays rank r of a w-way partition of its own rows (the same
 * count / scatter kernels; the output holds the rows rank r owns). */
#define PGF_XCHG_EMULATE(w, r) ((((uint32_t)(w)) & 0xFFu) << 8 | (((uint32_t)(r)) & 0xFFu) << 16)
pgf_status pgf_join_table_exchange(pgf_ctx *ctx, uint64_t table_or_rows, uint32_t mode, uint64_t *out_handle,
                                   uint64_t *nvlink_bytes_out);
/* rank that owns a key under PGF_XCHG_PARTITION (for tests) */
uint32_t pgf_partition_of_key(int64_t key, uint32_t world);

/* -------------------------------------------------- synthetic TPC-H-shaped data
 * Counter-based generator (row id -> values) writing reference-format pages directly in
 * HBM, so SF100 never exists on the host.  `table` selects the scan shape. */
enum {
  PGF_GEN_LINEITEM_Q6 = 1,  /* l_quantity, l_extendedprice, l_discount f64; l_shipdate utf8view */
  PGF_GEN_LINEITEM_Q1 = 2,  /* + l_tax f64; l_returnflag, l_linestatus, l_shipdate utf8view */
  PGF_GEN_LINEITEM_Q3 = 3,  /* l_orderkey i32; l_extendedprice, l_discount f64; l_shipdate */
  PGF_GEN_ORDERS_Q3 = 4,    /* o_orderkey, o_custkey i32; o_orderdate utf8view; o_shippriority i32 */
  PGF_GEN_CUSTOMER_Q3 = 5,  /* c_custkey i32; c_mktsegment utf8view */
  PGF_GEN_KEYS_I64 = 6,     /* one Int64 key column: key = splitmix64(seed + i) or i + 1 */
  /* D variants: the same rows with money as Decimal128(15,2), dates as Date32
   * (Int32 days since 1970-01-01) and flags as Int16 character codes */
  PGF_GEN_LINEITEM_Q6_D = 7, /* l_quantity, l_extendedprice, l_discount decimal; l_shipdate i32 */
  PGF_GEN_LINEITEM_Q1_D = 8  /* + l_tax decimal; l_returnflag, l_linestatus i16; l_shipdate i32 */
};
typedef struct {
  int32_t table;
  int32_t dense_keys;     /* PGF_GEN_KEYS_I64: 1 => key = first_row + i + 1 */
  uint64_t seed;
  uint64_t first_row;     /* global row id of this shard's first row */
  uint64_t rows;          /* rows to generate */
  uint64_t scale_rows;    /* table cardinality at this scale factor (key domains) */
} pgf_gen_spec;
/* Declares scan_id with the table's schema and fills it with generated pages. */
pgf_status pgf_gen_scan(pgf_ctx *ctx, uint64_t scan_id, const pgf_gen_spec *spec);
pgf_status pgf_gen_schema(int32_t table, pgf_column_spec *schema_out, uint32_t *ncols_out);

#ifdef __cplusplus
}
#endif
#endif /* PGF_B200_H */
