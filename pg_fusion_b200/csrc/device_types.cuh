// POD types shared between the host planner and the sm_100a kernels.
#pragma once
#include <cstdint>

namespace pgf {

constexpr uint32_t kMaxStageCols = 16;  // == PGF_MAX_COLS
constexpr uint32_t kKeyWords = 4;       // group key: up to 4 x u64 (32 bytes)
constexpr uint32_t kMaxExprs = 8;
constexpr uint32_t kMaxTerms = 8;
constexpr uint32_t kMaxJoins = 2;
constexpr uint32_t kMaxBlooms = 2;
constexpr uint32_t kRegGroups = 4;      // groups pre-aggregated in registers per CTA

// ---- HBM data layout of a scan ------------------------------------------------------
// pages: npages x page_size bytes, each the verbatim shared-memory page (20-byte transfer
// header + arrow_layout block), so every column buffer is 16-byte aligned physically.
struct PageDesc {        // 16 bytes per page, built and validated at ingest
  uint32_t row_count;
  uint16_t layout_class; // index into LayoutClass[] (pages with equal max_rows share one)
  uint16_t null_mask;    // bit c set: column c has null_count > 0 in this page
  uint64_t row_base;     // global row number of the page's first row
};
struct LayoutClass {     // byte offsets from the start of the PAGE (block offset + 20)
  uint32_t max_rows;
  uint32_t pool_base;
  uint32_t values_off[kMaxStageCols];
  uint32_t validity_off[kMaxStageCols];
};

// ---- Bloom filter parameters on the device -----------------------------------------
struct DevBloom {
  uint64_t* words;
  uint64_t bit_count;
  uint64_t seed;
  uint64_t m_lo, m_hi;   // ceil(2^128 / bit_count) for the exact 64-bit fastmod
  uint32_t hash_count;
  uint32_t pow2;         // 1: bit_count is a power of two (mask instead of modulo); 2: and in [32, 2^32]
};

// ---- pipeline plan -----------------------------------------------------------------
enum : uint8_t { SRC_PAGE = 0, kSrcRecord = 3 };  // 1, 2 = payload of join (src - 1); 3 = the scanned row of a row-set scan
enum : uint8_t { CLS_F64 = 0, CLS_I64 = 1, CLS_I128 = 2 };

struct DevStageCol {
  uint16_t page_col;     // column index in the page
  uint16_t width;        // bytes per row; 0 = bit-packed (Boolean): the tile holds tile_rows / 8 bytes
  uint32_t smem_off;     // offset of the column tile inside a stage
  uint32_t valid_off;    // offset of the validity-bitmap tile inside a stage (nullable only)
  uint16_t nullable;
  uint16_t type;
};

enum : uint8_t { LD_I16 = 0, LD_I32 = 1, LD_I64 = 2, LD_F32 = 3, LD_F64 = 4, LD_VIEW = 5, LD_DEC = 6,
                 LD_BOOL = 7 /* bit-packed values buffer (types.rs:139-147): predicates only */ };
constexpr uint32_t kNoValidity = 0xFFFFFFFFu;

struct DevRef {          // where a value comes from
  uint8_t src;           // SRC_PAGE or join index + 1
  uint8_t ld;            // LD_*: how to load / widen it
  uint8_t type;          // PGF_T_*
  uint8_t pcol;          // SRC_PAGE: page column index; else: payload column index
  uint32_t off;          // SRC_PAGE: byte offset of the column tile inside a stage; else first u32 payload word
  uint32_t valid_off;    // SRC_PAGE + nullable: offset of the validity tile; else kNoValidity
};

// One conjunct after host-side normalisation: every LT/LE/GT/GE/EQ on a column is folded
// into one inclusive range over the order-preserving key (signed k0, unsigned k1); NE
// becomes "not in [c, c]".
enum : uint32_t { TERM_IN_RANGE = 0, TERM_NOT_IN_RANGE = 1, TERM_NEVER = 2 };
struct DevTerm {
  DevRef ref;
  uint32_t op;           // TERM_*
  uint32_t wide;         // 0: one-word key, test (u64)(k0 - lo0) <= span; 1: (k0, k1) pair compare
  int64_t lo0;
  uint64_t lo1;          // wide: low word of the lower bound; narrow: span = hi0 - lo0
  int64_t hi0;
  uint64_t hi1;
};

struct DevFactor {
  DevRef ref;
  uint32_t kind;         // PGF_FACTOR_*
  uint32_t pad;
  double cf;             // constant (CLS_F64)
  int64_t ci_lo, ci_hi;  // constant (CLS_I64 / CLS_I128)
};
// form: straight-line fast paths for products of Float64 scan columns
enum : uint32_t { FORM_GENERIC = 0, FORM_X = 1, FORM_XY = 2, FORM_X_CMY = 3, FORM_X_CMY_CPZ = 4,
                  FORM_PREV_CPZ = 5 /* shapes only: previous argument times (c + z) */ };
struct DevExpr {
  uint32_t nfactors;
  uint32_t form;
  uint32_t null_cols;    // page columns (bit mask) whose validity this expression depends on
  uint32_t has_payload;  // some factor reads a join payload (may be NULL)
  DevFactor f[3];
};

struct DevKeyPart {
  DevRef ref;
  uint16_t word;         // first key word
  uint16_t nwords;       // 1 or 2
};

// Join table: open addressing with linear probing.  `tags` is a one-byte directory (0 = empty
// slot, else the top hash bits | 1) that stays L2 resident: probes walk the tags and touch a
// 16/32-byte slot only when the tag matches, so misses (most probes of a selective join)
// never read the slots.
struct DevJoin {
  const uint4* slots;    // slot_u4 x uint4 per slot: {key lo, key hi, occupied, pay0} [, pay1..4]
  const uint8_t* tags;   // capacity bytes
  uint32_t mask;         // capacity - 1
  uint32_t slot_u4;      // 1 (16-byte slot) or 2 (32-byte slot)
  DevRef key;
  uint32_t shift;        // 64 - log2(capacity): home bucket = (join_hash(key) >> shift) & ~7
  uint32_t pad;
};

// Join tables hash the sign-extended key with one 32 x 64-bit multiply (Fibonacci hashing): the top
// log2(capacity) bits pick the slot, rounded down to an 8-slot bucket whose tags are one aligned
// 8-byte word; the three dropped bits and the five below them make the tag byte (never 0 = empty).
constexpr uint32_t kJoinBucket = 8;
__host__ __device__ __forceinline__ uint64_t join_hash(int64_t key) {
  const uint32_t lo = uint32_t(uint64_t(key)), hi = uint32_t(uint64_t(key) >> 32);
  return uint64_t(lo ^ (hi * 0x85EBCA6Bu)) * 0x9E3779B97F4A7C15ull;
}
// owner rank of a key in a hash-partitioned exchange: bits of a second multiplicative hash, independent of the
// bits that place the key inside its owner's table
__host__ __device__ __forceinline__ uint32_t join_partition(int64_t key, uint32_t world) {
  const uint32_t lo = uint32_t(uint64_t(key)), hi = uint32_t(uint64_t(key) >> 32);
  uint64_t h = uint64_t(lo ^ (hi * 0x85EBCA6Bu)) * 0xD6E8FEB86659FD93ull;
  h ^= h >> 32;
  h *= 0xD6E8FEB86659FD93ull;
  return uint32_t(h >> 40) % world;
}
__host__ __device__ __forceinline__ uint32_t join_home(uint64_t h, uint32_t shift) { return uint32_t(h >> shift) & ~(kJoinBucket - 1u); }
__host__ __device__ __forceinline__ uint32_t join_tag8(uint64_t h, uint32_t shift) {
  const uint32_t b = uint32_t(h >> (shift - 5u)) & 0xFFu;
  return b ? b : 1u;
}

struct DevBloomProbe {
  DevBloom bloom;
  DevRef key;
};

// Global group table (open addressing, linear probing).  state: 0 empty, 1 locked, 2 ready;
// bits 8.. hold the key null mask.
struct GroupTable {
  uint32_t* state;
  uint64_t* keys;        // capacity x kKeyWords
  uint64_t* acc;         // capacity x nexprs x acc_words (f64 bits / i64 / i128 lo,hi)
  uint64_t* cnt;         // capacity x (nexprs + 1): per-expr non-null counts, then row count
  uint32_t mask;
  uint32_t acc_words;    // 1 or 2
  uint32_t nexprs;       // accumulators per slot (the thread that creates a group zeroes its acc / cnt words)
  uint32_t pad;
  uint32_t* overflow;    // set to 1 when the table is full
  uint32_t* used;        // number of occupied slots
};

struct JoinBuild {
  uint4* rows;             // dense build rows {key lo, key hi, occupancy/NULL flags, payload...}, slot_u4 x uint4 each
  uint64_t rows_cap;       // room in `rows`; the row count lives in the arena header (`used`)
  uint32_t slot_u4;
  uint32_t pad0;
  DevRef key;
  uint32_t npayload;
  DevRef payload[4];
  uint16_t payload_word[4];   // first u32 payload word of each payload column
  uint16_t payload_nwords[4];
};

struct Counters {        // RuntimeFilter*/Worker* style metrics, filled by the kernel
  unsigned long long rows_in, rows_bloom, rows_filtered, rows_out, bloom_rows, bad_rows;
};

enum : uint32_t { SINK_AGG = 1, SINK_JOIN_BUILD = 2, SINK_COUNT = 3 };

struct DevPlan {
  const uint8_t* pages;
  const PageDesc* descs;
  const LayoutClass* classes;
  uint64_t page_stride;
  uint32_t npages, tiles_per_page, tile_rows, nitems;
  uint32_t stage_bytes, nstage_cols;
  uint32_t nstages, pad0;   // ring depth: kStages partial-page tiles or 3 whole pages
  uint32_t used_null_mask;  // page columns whose validity matters to this plan
  uint32_t view_mask;       // page columns read as views (must be inline)
  DevStageCol scol[kMaxStageCols];

  uint32_t nbloom, nterms, njoins, sink;
  // compaction pipeline: 1 = bloom[0] is keyed on the column the queue entries carry (the first probe key, or its
  // own key when there is no join) and is probed AFTER the predicate, on dense lanes (stage B)
  uint32_t bloom_dense;
  // 1: every page of the scan has layout class 0, whose offsets travel in the plan itself (class0): stage C then
  // reads column offsets from the constant bank instead of chasing descs[page] -> classes[class] through HBM
  uint32_t single_class;
  uint32_t entry_key_off;   // single-term specialisation of the compaction pipeline: stage offset of the Int32 entry key
  // page columns stage C reads straight from HBM (late materialisation): it starts those reads before it walks the
  // probe chain, so the row's DRAM round trips overlap the chain's instead of following them
  uint32_t nlate;
  uint8_t late_pcol[8];
  uint8_t late_width[8];
  LayoutClass class0;
  DevBloomProbe bloom[kMaxBlooms];
  DevTerm terms[kMaxTerms];
  DevJoin joins[kMaxJoins];

  // SINK_AGG
  uint32_t nkeys, nkeywords, nexprs, acc_cls;
  DevKeyPart keys[4];
  DevExpr exprs[kMaxExprs];
  GroupTable table;
  // SINK_JOIN_BUILD
  JoinBuild build;
  DevBloom build_bloom;
  uint32_t has_build_bloom;
  uint32_t pad;
  unsigned long long* build_count;  // rows appended to build.rows so far (arena header)
  // Float64 aggregate sinks of the streaming kernel: per-CTA sums [grid][kRegGroups][2 + nexprs] and the
  // arrival counter of the fixed-order cross-CTA reduction
  uint64_t* cta_rec;
  uint32_t* cta_done;
  // Split execution of the compaction pipeline (joins): stages A and B append the rows whose tag window matched to
  // `entries` ({page, row, key lo, key hi}); a second, latency-oriented kernel runs stage C over them.
  uint4* entries;
  uint64_t entries_cap;
  unsigned long long* entries_count;   // entries appended (may exceed entries_cap: then *entries_overflow is set)
  uint32_t* entries_overflow;
  // row-set scans (rows_pipeline_kernel): dense rows in build-row format instead of pages
  const uint4* row_src;
  uint64_t row_count;
  uint32_t row_u4, pad1;

  Counters* counters;
};

}  // namespace pgf
