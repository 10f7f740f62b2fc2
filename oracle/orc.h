/*
 * oracle/orc.h -- CPU ORACLE for the pg_fusion worker-side columnar hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke() entry
 * point and bench.py's cpu_baseline / --impl reference legs may load it.  The
 * product library (pg_fusion_b200/csrc -> libpgf_b200.so) never links, imports
 * or calls anything in this directory.
 *
 * Every function restates, in plain C, the algorithm of the reference file:line
 * it cites (paths relative to the reference repository root).  Parts marked
 * [DF-K] restate DataFusion 44.0 / arrow-rs 53.4.1 semantics from knowledge of
 * those un-vendored crates (Cargo.toml:49-60); their source is not in the
 * reference tree, so for them parity is anchored on the reference's call sites
 * and on pyarrow/Acero cross-checks, and is "parity unpinned" at the DataFusion
 * operator boundary.  Bloom and page-layout parts are pinned by
 * the reference's own known-answer tests (tests/test_oracle_*.py).
 */
#ifndef PGF_ORACLE_H
#define PGF_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ Bloom */

/* runtime_filter/src/bloom.rs:17-22 */
typedef struct {
  uint64_t bit_count;
  uint64_t word_count;
  uint64_t hash_count;
  uint64_t seed;
} orc_bloom_params;

/* runtime_filter/src/bloom.rs:103-109 (BloomParamError) */
enum {
  ORC_OK = 0,
  ORC_BLOOM_ZERO_BIT_COUNT = 1,
  ORC_BLOOM_ZERO_HASH_COUNT = 2,
  ORC_BLOOM_ZERO_EXPECTED_ITEMS = 3,
  ORC_BLOOM_INVALID_FPR = 4,
  ORC_BLOOM_TOO_MANY_BITS = 5,
  ORC_BLOOM_INSUFFICIENT_WORDS = 6, /* BloomAttachError, bloom.rs:133-137 */
  ORC_BLOOM_NULL_BITS = 7
};

uint64_t orc_splitmix64(uint64_t v);                 /* bloom.rs:293-299 */
uint64_t orc_hash_int_key(int64_t v);                /* runtime_filter/src/lib.rs:31-34 */
int orc_bloom_params_new(uint64_t bit_count, uint64_t hash_count, uint64_t seed,
                         orc_bloom_params *out);     /* bloom.rs:29-48 */
int orc_bloom_params_for_expected_items(uint64_t expected_items, double fpr, uint64_t seed,
                                        orc_bloom_params *out); /* bloom.rs:52-79 */
int orc_bloom_attach_check(const orc_bloom_params *p, const uint64_t *bits,
                           uint64_t words_available); /* bloom.rs:167-178 */
uint64_t orc_bloom_bit_index(const orc_bloom_params *p, uint64_t hash,
                             uint64_t hash_index);   /* bloom.rs:250-255 */
void orc_bloom_clear(const orc_bloom_params *p, uint64_t *bits);             /* bloom.rs:205-209 */
void orc_bloom_insert_hash(const orc_bloom_params *p, uint64_t *bits, uint64_t hash); /* :222-227 */
int orc_bloom_might_contain_hash(const orc_bloom_params *p, const uint64_t *bits,
                                 uint64_t hash);     /* bloom.rs:233-241 */

/* Batch drivers.  key_width in {2,4,8}: Int16/Int32/Int64 keys are sign-extended to
 * i64 then cast to u64 (worker_runtime/src/runtime_filter_plan.rs:244,256,268;
 * pg/slot_encoder/src/encoder.rs:357-365).  validity: LSB-first bitmap or NULL.
 * insert: non-null rows only (runtime_filter_plan.rs:345-363); returns rows inserted.
 * probe: keep[i]=1 unless DefinitelyAbsent; NULL key => DefinitelyAbsent when the
 * filter is Ready (runtime_filter/src/shared.rs:350-374;
 * pg/backend_service/src/source.rs:496-532).  Returns rows rejected. */
uint64_t orc_bloom_insert_keys(const orc_bloom_params *p, uint64_t *bits, const void *keys,
                               int key_width, const uint8_t *validity, uint64_t n);
uint64_t orc_bloom_probe_keys(const orc_bloom_params *p, const uint64_t *bits, const void *keys,
                              int key_width, const uint8_t *validity, uint64_t n, uint8_t *keep);

/* Lifecycle word: runtime_filter/src/shared.rs:7-9,400-416 */
enum { ORC_RF_FREE = 0, ORC_RF_BUILDING = 1, ORC_RF_READY = 2, ORC_RF_DISABLED = 3 };
enum { ORC_PASS_UNFILTERED = 0, ORC_MAYBE_PRESENT = 1, ORC_DEFINITELY_ABSENT = 2 };
enum {
  ORC_LC_OK = 0,
  ORC_LC_GENERATION_EXHAUSTED = 1,
  ORC_LC_BUSY = 2,
  ORC_LC_INVALID_TRANSITION = 3
};
int orc_lifecycle_pack(uint64_t generation, int state, uint64_t *word); /* shared.rs:400-408 */
void orc_lifecycle_unpack(uint64_t word, uint64_t *generation, int *state); /* shared.rs:411-416 */
/* shared.rs:159-198: Free|Disabled -> (gen+1, Building), clears bits. */
int orc_slot_try_acquire_builder(uint64_t *lifecycle, const orc_bloom_params *p, uint64_t *bits,
                                 uint64_t *generation_out);
int orc_slot_publish_build(uint64_t *lifecycle, uint64_t generation);   /* shared.rs:201-208,377-397 */
int orc_slot_disable_build(uint64_t *lifecycle, uint64_t generation);   /* shared.rs:211-213 */
int orc_slot_retire_ready(uint64_t *lifecycle, uint64_t generation);    /* shared.rs:244-260 */
int orc_probe_decision_for_hash(const uint64_t *lifecycle, uint64_t generation,
                                const orc_bloom_params *p, const uint64_t *bits,
                                uint64_t hash);                         /* shared.rs:350-361 */
int orc_probe_decision_for_null(const uint64_t *lifecycle, uint64_t generation); /* :367-374 */

/* ------------------------------------------------------------ Page layout */

/* page/arrow_layout/src/constants.rs:4-26 */
#define ORC_BLOCK_MAGIC 0x32424150u
#define ORC_BLOCK_VERSION 1u
#define ORC_BUFFER_ALIGNMENT 16u
#define ORC_BUFFER_ALIGNMENT_BIAS 12u
#define ORC_VIEW_INLINE_LEN 12u
/* page/transfer/src/page.rs:8-11; page/import/src/lib.rs:43 */
#define ORC_PAGE_MAGIC 0x50545031u
#define ORC_PAGE_HEADER_LEN 20u
#define ORC_ARROW_LAYOUT_BATCH_KIND 0x4152u

/* page/arrow_layout/src/types.rs:93-112 */
enum {
  ORC_T_BOOLEAN = 1,
  ORC_T_INT16 = 2,
  ORC_T_INT32 = 3,
  ORC_T_INT64 = 4,
  ORC_T_FLOAT32 = 5,
  ORC_T_FLOAT64 = 6,
  ORC_T_UUID = 7,
  ORC_T_UTF8VIEW = 8,
  ORC_T_BINARYVIEW = 9,
  /* EXTENSION beyond reference v1 (SURVEY 8d "D" schema): Decimal128 as a 16-byte
   * fixed-width little-endian two's-complement slot.  Not accepted by the reference. */
  ORC_T_DECIMAL128 = 10
};
#define ORC_COLFLAG_NULLABLE 1u /* types.rs:44 */
#define ORC_COLFLAG_VIEW 2u     /* types.rs:46 */

/* page/arrow_layout/src/raw.rs:21-46 (40 bytes, align 4) */
typedef struct {
  uint32_t magic;
  uint16_t version;
  uint16_t flags;
  uint32_t block_size;
  uint32_t max_rows;
  uint32_t row_count;
  uint16_t col_count;
  uint16_t reserved0;
  uint32_t front_base;
  uint32_t pool_base;
  uint32_t tail_cursor;
  uint32_t reserved1;
} orc_block_header;

/* page/arrow_layout/src/raw.rs:69-84 (20 bytes, align 4) */
typedef struct {
  uint16_t type_tag;
  uint16_t flags;
  uint32_t validity_off;
  uint32_t values_off;
  uint32_t null_count;
  uint32_t reserved0;
} orc_column_desc;

/* page/arrow_layout/src/raw.rs:106-110 (16 bytes, align 4) */
typedef struct {
  int32_t len;
  uint8_t data[12];
} orc_byte_view;

typedef struct {
  uint16_t type_tag;
  uint16_t nullable;
} orc_column_spec; /* types.rs:218-225 */

typedef struct {
  uint16_t type_tag;
  uint16_t flags;
  uint32_t validity_off;
  uint32_t values_off;
  uint32_t validity_len;
  uint32_t values_len;
} orc_column_layout; /* types.rs:250-264 */

#define ORC_MAX_COLS 64
typedef struct {
  uint32_t block_size;
  uint32_t max_rows;
  uint32_t front_base;
  uint32_t pool_base;
  uint32_t ncols;
  orc_column_layout cols[ORC_MAX_COLS];
} orc_layout_plan; /* plan.rs:22-29 */

/* LayoutError / ImportError variants (page/arrow_layout/src/error.rs,
 * page/import/src/error.rs), numbered for the tests. */
enum {
  ORC_LE_INVALID_MAGIC = 101,
  ORC_LE_INVALID_VERSION = 102,
  ORC_LE_ROW_COUNT_EXCEEDS_MAX_ROWS = 103,
  ORC_LE_COLUMN_COUNT_MISMATCH = 104,
  ORC_LE_FRONT_BASE_MISMATCH = 105,
  ORC_LE_INVALID_HEADER_BOUNDS = 106,
  ORC_LE_MISALIGNED_FRONT_REGION = 107,
  ORC_LE_BLOCK_SLICE_TOO_SMALL = 108,
  ORC_LE_INVALID_TYPE_TAG = 109,
  ORC_LE_INCONSISTENT_VIEW_FLAG = 110,
  ORC_LE_COLUMN_DESC_MISMATCH = 111,
  ORC_LE_POOL_BASE_MISMATCH = 112,
  ORC_LE_LAYOUT_DOES_NOT_FIT = 113,
  ORC_LE_SIZE_OVERFLOW = 114,
  ORC_LE_TOO_MANY_COLUMNS = 115,
  ORC_LE_INLINE_VALUE_TOO_LARGE = 116,
  ORC_LE_NEGATIVE_VIEW_LENGTH = 117,
  ORC_LE_INVALID_VIEW_BUFFER_INDEX = 118,
  ORC_LE_NEGATIVE_VIEW_OFFSET = 119,
  ORC_LE_VIEW_OFFSET_OUT_OF_BOUNDS = 120,
  ORC_LE_COLUMN_INDEX_OUT_OF_BOUNDS = 121,
  ORC_LE_VIEW_FULL = 122, /* ViewWriteStatus::Full, access.rs:15-20 */
  ORC_IE_WRONG_KIND = 201,
  ORC_IE_UNSUPPORTED_FLAGS = 202,
  ORC_IE_SCHEMA_COLUMN_COUNT_MISMATCH = 203,
  ORC_IE_SCHEMA_TYPE_MISMATCH = 204,
  ORC_IE_SCHEMA_NULLABILITY_MISMATCH = 205,
  ORC_IE_INVALID_NULL_COUNT = 206,
  ORC_IE_NULL_BITMAP_COUNT_MISMATCH = 207,
  ORC_IE_VIEW_OFFSET_BEFORE_ALLOCATED_TAIL = 208,
  ORC_IE_PAGE_HEADER_INVALID = 209,
  /* arrow-rs StringViewArray/BinaryViewArray::try_new view validation [DF-K]:
   * non-zero inline padding, prefix mismatch, invalid UTF-8 (page/import/src/lib.rs:362-366,384-388;
   * pinned by page/import/src/tests.rs:508-528) */
  ORC_IE_ARROW_INVALID_VIEW = 210
};

int orc_type_row_width(int type_tag);                       /* types.rs:139-147 */
int orc_layout_plan_new(const orc_column_spec *specs, uint32_t ncols, uint32_t max_rows,
                        uint32_t block_size, orc_layout_plan *out); /* plan.rs:33-93 */
/* page/row_estimator/src/lib.rs:353-371 */
int orc_fixed_row_cap(const orc_column_spec *specs, uint32_t ncols, uint32_t block_size,
                      uint32_t *cap_out);
int orc_init_block(uint8_t *block, size_t len, const orc_layout_plan *plan); /* access.rs:640-654 */
int orc_block_validate(const uint8_t *block, size_t len);   /* BlockRef::open, access.rs:36-42 (+ the Decimal128 extension tag) */
int orc_block_validate_v1(const uint8_t *block, size_t len); /* the same, reference v1 tags 1..9 only */
int orc_block_write_fixed(uint8_t *block, size_t len, uint32_t col, uint32_t row,
                          const void *bytes, uint32_t nbytes); /* access.rs:316-319 */
int orc_block_write_bool(uint8_t *block, size_t len, uint32_t col, uint32_t row, int value);
int orc_block_write_null(uint8_t *block, size_t len, uint32_t col, uint32_t row); /* :322-337 */
int orc_block_write_view_bytes(uint8_t *block, size_t len, uint32_t col, uint32_t row,
                               const void *bytes, uint32_t nbytes); /* access.rs:341-366 */
int orc_block_commit_current_row(uint8_t *block, size_t len);      /* access.rs:443-457 */
int orc_block_set_validity(uint8_t *block, size_t len, uint32_t col, uint32_t row, int valid);

/* transfer page header (msgpack array of 5), page/transfer/src/page.rs:20-64,66-126 */
int orc_page_header_encode(uint16_t kind, uint16_t flags, uint32_t payload_len, uint8_t out[20]);
int orc_page_header_decode(const uint8_t in[20], uint16_t *kind, uint16_t *flags,
                           uint32_t *payload_len);

/* page/import/src/lib.rs:117-206 (+208-235, 237-293, 424-452): all import-time checks. */
int orc_import_check(uint16_t kind, uint16_t flags, const uint8_t *block, size_t len,
                     const orc_column_spec *schema, uint32_t ncols);

/* ------------------------------------------------- Relational operators [DF-K] */

/* A decoded column: the concatenation, in page order, of one column of every page of
 * a scan (what the DataFusion operators see as a stream of RecordBatches). */
typedef struct {
  uint32_t len;
  uint32_t pad;
  const uint8_t *ptr; /* points into the owning column's arena */
} orc_str;

typedef struct {
  int32_t type_tag;
  int32_t nullable;
  uint64_t rows;
  void *values;      /* fixed width: rows * width bytes; Boolean: byte per row;
                        Utf8View/BinaryView: rows * orc_str (inline and long values) */
  uint8_t *validity; /* byte per row (1 = valid), or NULL when no nulls */
  uint8_t *arena;    /* string bytes for view columns, else NULL */
} orc_column;

typedef struct {
  uint32_t ncols;
  uint64_t rows;
  orc_column cols[ORC_MAX_COLS];
} orc_table;

/* Decode a run of whole transfer pages (page_stride bytes apart, each a 20-byte
 * transfer header followed by one arrow_layout block) into a table; runs
 * orc_import_check on each page first. */
int orc_table_from_pages(const uint8_t *pages, uint64_t npages, uint64_t page_stride,
                         const orc_column_spec *schema, uint32_t ncols, orc_table *out);
void orc_table_free(orc_table *t);
/* filter_record_batch / take [DF-K]: out = rows of `in` with keep[i] != 0 / rows[i]. */
int orc_table_select(const orc_table *in, const uint8_t *keep, orc_table *out);
int orc_table_take(const orc_table *in, const uint64_t *rows, uint64_t n, orc_table *out);

/* Expression tree in postfix order (evaluated with a value stack).  Semantics [DF-K]:
 * arithmetic per row exactly as written, IEEE Float64 with no FMA contraction, integer
 * ops wrapping; comparisons yield SQL three-valued booleans; AND is Kleene; a row
 * passes a filter iff the predicate is TRUE. */
enum {
  ORC_X_COL = 1,     /* a = column index (of current row source), b = source (0 = probe/scan side, 1.. = join build side k) */
  ORC_X_LIT_F64 = 2, /* f */
  ORC_X_LIT_I64 = 3, /* i */
  ORC_X_LIT_STR = 4, /* s, slen (<= 12) */
  ORC_X_LIT_I128 = 5,/* i (lo), i2 (hi) */
  ORC_X_ADD = 10,
  ORC_X_SUB = 11,
  ORC_X_MUL = 12,
  ORC_X_LT = 20,
  ORC_X_LE = 21,
  ORC_X_GT = 22,
  ORC_X_GE = 23,
  ORC_X_EQ = 24,
  ORC_X_NE = 25,
  ORC_X_AND = 30
};
typedef struct {
  int32_t op;
  int32_t a;
  int32_t b;
  int32_t slen;
  double f;
  int64_t i;
  int64_t i2;
  char s[16];
} orc_xnode;

enum { ORC_AGG_SUM = 1, ORC_AGG_AVG = 2, ORC_AGG_COUNT_STAR = 3, ORC_AGG_COUNT = 4, ORC_AGG_MIN = 5, ORC_AGG_MAX = 6 };
typedef struct {
  int32_t func;
  int32_t expr_off; /* offset of the argument expression in the node array */
  int32_t expr_len; /* 0 for COUNT(*) */
} orc_agg_spec;

/* Value kinds in results */
enum { ORC_V_NULL = 0, ORC_V_F64 = 1, ORC_V_I64 = 2, ORC_V_I128 = 3, ORC_V_STR = 4, ORC_V_BOOL = 5 };
typedef struct {
  int32_t kind;
  int32_t slen;
  double f;
  int64_t lo; /* i64 value, or low 64 bits of i128 */
  int64_t hi; /* high 64 bits of i128 */
  char s[16];
} orc_value;

/* One inner equi-join edge of a left-deep pipeline: the probe-side row source is
 * source 0 (the scan) — key expression is column probe_col of source probe_src —
 * and the build side is a table with an optional filter (postfix nodes over its own
 * columns) keyed by build_col.  NULL keys never match; duplicates multiply. */
typedef struct {
  const orc_table *build;
  int32_t build_col;
  int32_t probe_src;
  int32_t probe_col;
} orc_join_edge;

typedef struct {
  uint64_t ngroups;
  uint32_t nkeys;
  uint32_t naggs;
  orc_value *keys; /* ngroups * nkeys, first-appearance order */
  orc_value *aggs; /* ngroups * naggs */
  uint64_t rows_in;
  uint64_t rows_filtered; /* rows passing the filter (before joins) */
  uint64_t rows_joined;   /* rows reaching the aggregate */
} orc_agg_result;

/* FilterExec -> HashJoinExec* -> AggregateExec(mode=Single) over `scan`.  [DF-K]
 * filter: postfix predicate over source 0, may be NULL/0.  joins: njoins edges applied
 * in order; source k+1 is edge k's build table (already filtered by the caller).
 * group keys: postfix expressions (usually single ORC_X_COL nodes).  Float64 SUM/AVG
 * per group are accumulated strictly sequentially in input row order (the grouped
 * accumulator path); with nkeys == 0 and sum_lanes > 0 the no-group path sums each
 * batch of batch_rows filtered rows in `sum_lanes` striped lanes, folds lanes pairwise,
 * and adds batch partials sequentially (arrow `sum` kernel shape). */
int orc_aggregate(const orc_table *scan, const orc_xnode *nodes, int32_t filter_off,
                  int32_t filter_len, const orc_join_edge *joins, uint32_t njoins,
                  const int32_t *key_off, const int32_t *key_len, uint32_t nkeys,
                  const orc_agg_spec *aggs, uint32_t naggs, int32_t sum_lanes,
                  int32_t batch_rows, orc_agg_result *out);
void orc_agg_result_free(orc_agg_result *r);

/* FilterExec alone: keep[i] = 1 iff predicate TRUE.  Returns kept rows via *kept. */
int orc_filter(const orc_table *scan, const orc_xnode *nodes, int32_t filter_off,
               int32_t filter_len, uint8_t *keep, uint64_t *kept);

/* HashJoinExec(CollectLeft, Inner) row-pair multiset: emits (build_row, probe_row)
 * pairs; returns count; pairs may be NULL to only count.  Keys Int16/32/64. */
int orc_hash_join_pairs(const orc_table *build, int32_t build_col, const orc_table *probe,
                        int32_t probe_col, uint64_t *build_rows, uint64_t *probe_rows,
                        uint64_t cap, uint64_t *npairs);

/* --------------------------------------------------- Fast CPU baselines (bench) */
/* Tight, multi-threadable loops for the three TPC-H shapes over whole transfer pages
 * of the reference-faithful "F" schema.  Checked against orc_aggregate in
 * tests.  nthreads >= 1: pages are range-sharded over threads and partial states are
 * merged in thread order (Partial -> Final). */
typedef struct {
  double sum;
  uint64_t rows_in;
  uint64_t rows_kept;
} orc_q6_result;
/* cols = {quantity, extendedprice, discount, shipdate} column indices in the page */
int orc_q6_pages(const uint8_t *pages, uint64_t npages, uint64_t page_stride, int nthreads,
                 const int32_t cols[4], const char *date_lo, const char *date_hi,
                 double disc_lo, double disc_hi, double qty_lt, orc_q6_result *out);

typedef struct {
  char returnflag;
  char linestatus;
  double sum_qty, sum_base_price, sum_disc_price, sum_charge, sum_disc;
  uint64_t count;
} orc_q1_group;
typedef struct {
  uint32_t ngroups;
  orc_q1_group groups[16];
  uint64_t rows_in;
} orc_q1_result;
/* cols = {quantity, extendedprice, discount, tax, returnflag, linestatus, shipdate} */
int orc_q1_pages(const uint8_t *pages, uint64_t npages, uint64_t page_stride, int nthreads,
                 const int32_t cols[7], const char *date_le, int with_tax, orc_q1_result *out);

#ifdef __cplusplus
}
#endif

/* ---- orc_q3.c: TPC-H Q3 shape, page-sharded, fed shard by shard (test / bench baseline only) ---- */
typedef struct orc_q3 orc_q3;
orc_q3 *orc_q3_new(const char *segment, const char *date);
void orc_q3_free(orc_q3 *q);
int orc_q3_customer(orc_q3 *q, const uint8_t *pages, uint64_t npages, uint64_t stride, int nthreads);
int orc_q3_customer_finish(orc_q3 *q);
int orc_q3_orders(orc_q3 *q, const uint8_t *pages, uint64_t npages, uint64_t stride, int nthreads);
int orc_q3_orders_finish(orc_q3 *q, int nthreads);
int orc_q3_lineitem(orc_q3 *q, const uint8_t *pages, uint64_t npages, uint64_t stride, int nthreads);
int orc_q3_stats(const orc_q3 *q, uint64_t out[6]);
uint64_t orc_q3_groups(const orc_q3 *q, int32_t *keys, uint8_t *dates12, int32_t *datelens, int32_t *prios, double *sums,
                       uint64_t *counts, uint64_t cap);

#endif
