// The fused scan pipeline kernel (sm_100a).
//
// One persistent CTA per SM streams row tiles of the scan's 64 KiB pages from HBM into a
// 4-stage shared-memory ring with TMA bulk copies (cp.async.bulk + mbarrier transaction
// counts; SASS: UBLKCP), issued by one producer warp.  16 consumer warps evaluate, per
// row and entirely in registers:
//   runtime Bloom probe(s)  -> K2  (pg/backend_service/src/source.rs:496-532)
//   FilterExec predicate     -> K3  (conjunction of <column> <cmp> <literal>)
//   ProjectionExec exprs     -> K4  (products of x, (c - x), (c + x))
//   sink: AggregateExec      -> K5  (no-group / register pre-aggregated / global hash)
// so every page byte is read from HBM exactly once and nothing is materialised.  Plans with a
// HashJoinExec probe or a build sink (K6 / K1) run the compaction pipeline in probe_kernel.cuh.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <type_traits>

#include "../../include/pgf_b200.h"
#include "bloom_device.cuh"
#include "device_types.cuh"

namespace pgf {

// Consumer warps per CTA: 16 for streaming sinks; 8 when kRegGroups x kMaxExprs register
// accumulators per thread are live (GROUP BY), which needs the larger register budget (12 warps).
constexpr int kMaxConsumerWarps = 16;
// `fast`: the fast GROUP BY path keeps its accumulators in shared memory, so it fits 16 warps too.
// The Decimal128 flavour of the fast path runs 15 + 1 warps: 512 threads may use 128 registers each (ptxas caps a
// block of 513..640 threads at 96), and its two rows of five 128-bit products spilled under 96.
__host__ __device__ constexpr int consumer_warps(uint32_t sink, bool grouped, bool fast = false, uint32_t acc = CLS_F64) {
  return (sink == SINK_AGG && grouped && !fast) ? 14 : ((fast && acc == CLS_I128) ? 15 : 16);
}
// Producer warps: one feeds the ring of most shapes; the fast GROUP BY path (Q1: 80 bytes per row, seven bulk copies per
// tile) was starved by a single one (ncu r1: 44 % of its stall samples on the `full` barrier), so there two producers
// alternate tiles.
__host__ __device__ constexpr int producer_warps(bool fast = false, uint32_t acc = CLS_F64) { return (fast && acc != CLS_I128) ? 2 : 1; }
__host__ __device__ constexpr int pipeline_threads(uint32_t sink, bool grouped, bool fast = false, uint32_t acc = CLS_F64) {
  return (consumer_warps(sink, grouped, fast, acc) + producer_warps(fast, acc)) * 32;
}
constexpr int kStages = 4;     // ring depth of streaming pipelines
#ifndef PGF_JOIN_ROWS
#define PGF_JOIN_ROWS 4
#endif
// The ring depth is a plan parameter (DevPlan::nstages): kStages tiles of a fraction of a page,
// or 3 whole pages -- whichever keeps more consumer warps busy per tile (see layout_stage()).
__host__ __device__ constexpr uint32_t rows_per_thread(uint32_t nj) { return nj != 0 ? PGF_JOIN_ROWS : 2; }
// Deferred sink: rows whose group is not register resident are queued per warp (in shared
// memory behind the stage ring) and applied to the global table 32 at a time from a converged
// point, instead of one or two lanes at a time from inside the divergent probe loop.
constexpr uint32_t kQueueBytesPerWarp = 2048;
__host__ __device__ constexpr uint32_t queue_entry_words(uint32_t maxe, uint32_t acc_words) { return kKeyWords + 1 + maxe * acc_words; }
// shared-memory accumulator slots of the fast GROUP BY path: groups x (arguments + row count) x consumer threads x 8 B
__host__ __device__ constexpr uint32_t fast_group_acc_bytes(uint32_t nexprs) { return kRegGroups * (nexprs + 1) * uint32_t(consumer_warps(SINK_AGG, true, true)) * 32u * 8u; }
constexpr uint32_t kAccF64MaxExprs = 8, kAccI128MaxExprs = 6;

struct StageMeta {
  uint32_t nrows;
  uint32_t null_mask;
  uint64_t row_base;
};

// ---- compile-time plan shapes ----------------------------------------------------------
// The kernel is written once, generically over the plan grammar (every per-term / per-
// expression decision is a warp-uniform runtime dispatch).  For the plan shapes that matter
// most, the same code is instantiated with the load kinds of the predicate terms and the
// forms of the aggregate arguments as compile-time constants, which folds the dispatches
// away.  Literals, column positions and bounds stay runtime parameters in both cases.
template <int... V>
struct IntList {
  static constexpr int size = int(sizeof...(V));
  template <int I>
  static constexpr int at() {
    constexpr int a[] = {V..., -1};
    return I < size ? a[I < size ? I : 0] : -1;
  }
};
// Key part encoding for shapes: LD_* | (1 << 4 if the part is a join payload) | (first key word << 8)
constexpr int key_enc(int ld, bool payload, int word) { return ld | (payload ? 16 : 0) | (word << 8); }

template <bool GENERIC, class TERMS, class EXPRS, bool NONULL = false, class KEYS = IntList<>>
struct ShapeT {
  using Keys = KEYS;     // key_enc() of every GROUP BY key part (empty: resolved at run time)
  static constexpr bool generic = GENERIC;
  static constexpr bool no_nulls = NONULL;  // no nullable scan column is referenced: validity checks compile out
  using Terms = TERMS;   // LD_* of every (range) predicate term, in plan order
  using Exprs = EXPRS;   // FORM_* of every aggregate argument
};
using GenericShape = ShapeT<true, IntList<>, IntList<>>;

template <int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (N > 0) {
    static_for<N - 1>(f);
    f(std::integral_constant<int, N - 1>{});
  }
}

// ---- PTX wrappers: mbarrier + TMA bulk copy -----------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "PGF_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra PGF_DONE;\n"
      "bra PGF_WAIT;\n"
      "PGF_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- small helpers ------------------------------------------------------------------
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  return x;
}
__device__ __forceinline__ uint32_t bswap32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }

struct I128 {
  uint64_t lo, hi;
};
__device__ __forceinline__ I128 i128_add(I128 a, I128 b) {
  I128 r;
  r.lo = a.lo + b.lo;
  r.hi = a.hi + b.hi + (r.lo < a.lo ? 1ull : 0ull);
  return r;
}
__device__ __forceinline__ I128 i128_sub(I128 a, I128 b) {
  I128 r;
  r.lo = a.lo - b.lo;
  r.hi = a.hi - b.hi - (a.lo < b.lo ? 1ull : 0ull);
  return r;
}
__device__ __forceinline__ I128 i128_mul(I128 a, I128 b) {  // low 128 bits (wrapping)
  I128 r;
  r.lo = a.lo * b.lo;
  r.hi = __umul64hi(a.lo, b.lo) + a.lo * b.hi + a.hi * b.lo;
  return r;
}

// The same low 128 bits, cheaper when the operands are narrow -- which money columns are: a Decimal128 that is the
// sign extension of its low 32 (64) bits multiplies as a signed 32 x 32 -> 64 (64 x 64 -> 128) product, and the
// wrapping product of two sign-extended numbers IS the sign extension of that product.  One IMAD.WIDE against ~24
// integer instructions for the general case; the branch is taken by whole warps in practice.
__device__ __forceinline__ I128 i128_mul_narrow(I128 a, I128 b) {
  const int32_t a0 = int32_t(uint32_t(a.lo)), b0 = int32_t(uint32_t(b.lo));
  const bool a32 = int64_t(a.lo) == int64_t(a0) && a.hi == uint64_t(int64_t(a0) >> 63);
  const bool b32 = int64_t(b.lo) == int64_t(b0) && b.hi == uint64_t(int64_t(b0) >> 63);
  if (a32 && b32) {
    const int64_t p = int64_t(a0) * int64_t(b0);
    return I128{uint64_t(p), uint64_t(p >> 63)};
  }
  if (a.hi == uint64_t(int64_t(a.lo) >> 63) && b.hi == uint64_t(int64_t(b.lo) >> 63))
    return I128{a.lo * b.lo, uint64_t(__mul64hi((long long)a.lo, (long long)b.lo))};
  return i128_mul(a, b);
}

template <uint32_t ACC>
struct AccOps;
template <>
struct AccOps<CLS_F64> {
  using T = double;
  static constexpr uint32_t kMaxExprs = kAccF64MaxExprs;
  static __device__ __forceinline__ T zero() { return 0.0; }
  static __device__ __forceinline__ T add(T a, T b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ T shfl_xor(T v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
  static __device__ __forceinline__ void atomic_add(uint64_t* p, T v) { atomicAdd(reinterpret_cast<double*>(p), v); }
};
template <>
struct AccOps<CLS_I64> {
  using T = uint64_t;
  static constexpr uint32_t kMaxExprs = kAccF64MaxExprs;
  static __device__ __forceinline__ T zero() { return 0; }
  static __device__ __forceinline__ T add(T a, T b) { return a + b; }
  static __device__ __forceinline__ T shfl_xor(T v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
  static __device__ __forceinline__ void atomic_add(uint64_t* p, T v) {
    atomicAdd(reinterpret_cast<unsigned long long*>(p), static_cast<unsigned long long>(v));
  }
};
template <>
struct AccOps<CLS_I128> {
  using T = I128;
  static constexpr uint32_t kMaxExprs = kAccI128MaxExprs;
  static __device__ __forceinline__ T zero() { return I128{0, 0}; }
  static __device__ __forceinline__ T add(T a, T b) { return i128_add(a, b); }
  static __device__ __forceinline__ T shfl_xor(T v, int o) {
    return I128{__shfl_xor_sync(0xffffffffu, v.lo, o), __shfl_xor_sync(0xffffffffu, v.hi, o)};
  }
  static __device__ __forceinline__ void atomic_add(uint64_t* p, T v) {
    // 128-bit wrapping add as two 64-bit atomics with carry propagation (commutative)
    const unsigned long long old = atomicAdd(reinterpret_cast<unsigned long long*>(p), static_cast<unsigned long long>(v.lo));
    const unsigned long long carry = (old + v.lo) < old ? 1ull : 0ull;
    atomicAdd(reinterpret_cast<unsigned long long*>(p + 1), static_cast<unsigned long long>(v.hi + carry));
  }
};

// ---- per-row access ---------------------------------------------------------------------
// A row of the staged tile plus (behind a join probe) the matched build-side slot.
struct Row {
  const uint8_t* stage;
  uint32_t r;
  uint32_t tile_nulls;   // page columns with nulls in this tile (already masked by use)
  const uint32_t* pay;   // matched join slot (u32 words: key lo, key hi, occ, payload...)
  uint32_t occ;          // slot occupancy word: bit 0 occupied, bit 1+p payload p is NULL
};

__device__ __forceinline__ bool ref_valid(const DevRef& ref, const Row& row) {
  if (ref.src == SRC_PAGE) {
    if (ref.valid_off == kNoValidity || !((row.tile_nulls >> ref.pcol) & 1)) return true;
    return (row.stage[ref.valid_off + (row.r >> 3)] >> (row.r & 7)) & 1;
  }
  return !((row.occ >> (1 + ref.pcol)) & 1);
}

__device__ __forceinline__ int64_t load_i64(const DevRef& ref, const Row& row) {
  if (ref.src == SRC_PAGE) {
    const uint8_t* p = row.stage + ref.off;
    switch (ref.ld) {
      case LD_I16: return int64_t(reinterpret_cast<const int16_t*>(p)[row.r]);
      case LD_I32: return int64_t(reinterpret_cast<const int32_t*>(p)[row.r]);
      default: return reinterpret_cast<const int64_t*>(p)[row.r];
    }
  }
  const uint32_t* p = row.pay + 3 + ref.off;
  switch (ref.ld) {
    case LD_I16: return int64_t(int16_t(__ldg(p)));
    case LD_I32: return int64_t(int32_t(__ldg(p)));
    default: return int64_t((uint64_t(__ldg(p + 1)) << 32) | __ldg(p));
  }
}

__device__ __forceinline__ uint4 load_u128(const DevRef& ref, const Row& row) {
  if (ref.src == SRC_PAGE) return reinterpret_cast<const uint4*>(row.stage + ref.off)[row.r];
  const uint32_t* p = row.pay + 3 + ref.off;
  return make_uint4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
}

__device__ __forceinline__ double load_f64(const DevRef& ref, const Row& row) {
  if (ref.src == SRC_PAGE) {
    const uint8_t* p = row.stage + ref.off;
    switch (ref.ld) {
      case LD_F64: return reinterpret_cast<const double*>(p)[row.r];
      case LD_F32: return double(reinterpret_cast<const float*>(p)[row.r]);
      default: return double(load_i64(ref, row));  // AVG over integers runs on the Float64 cast
    }
  }
  const uint32_t* p = row.pay + 3 + ref.off;
  if (ref.ld == LD_F64) return __longlong_as_double((long long)((uint64_t(__ldg(p + 1)) << 32) | __ldg(p)));
  if (ref.ld == LD_F32) return double(__uint_as_float(__ldg(p)));
  return double(load_i64(ref, row));
}

// ---- FilterExec: one normalised conjunct -------------------------------------------------
// Keys are order preserving: integers as themselves, floats by IEEE totalOrder (what arrow's
// comparison kernels use), inline views (<= 12 bytes, zero padded) as (big-endian first 8
// bytes, big-endian next 4 bytes, length), decimals as (hi, lo).
__device__ __forceinline__ int64_t f64_key(int64_t bits) {
  // totalOrder key on 32-bit halves: 3 instructions instead of 64-bit shifts
  const int32_t hi = int32_t(uint64_t(bits) >> 32);
  const int32_t m = hi >> 31;
  const uint32_t klo = uint32_t(uint64_t(bits)) ^ uint32_t(m);
  const uint32_t khi = uint32_t(hi) ^ (uint32_t(m) >> 1);
  return int64_t((uint64_t(khi) << 32) | klo);
}

__device__ __forceinline__ bool in_range1(int64_t k, const DevTerm& T) { return uint64_t(k - T.lo0) <= T.lo1; }
// Wide keys are unsigned 128-bit (hi, lo); bounds are stored as lo = (lo0:lo1), span = (hi0:hi1).
__device__ __forceinline__ bool in_range2(uint64_t khi, uint64_t klo, const DevTerm& T) {
  const unsigned __int128 k = (static_cast<unsigned __int128>(khi) << 64) | klo;
  const unsigned __int128 lo = (static_cast<unsigned __int128>(uint64_t(T.lo0)) << 64) | T.lo1;
  const unsigned __int128 span = (static_cast<unsigned __int128>(uint64_t(T.hi0)) << 64) | T.hi1;
  return (k - lo) <= span;
}
// An out-of-line view (length > 12, page/arrow_layout/src/raw.rs:98-110) keeps bytes 0..3 in the view and the whole
// value in the page's tail arena at `offset` (buffer index, bounds and prefix were verified by the import checks,
// ingest.cu).  The order-preserving key of a predicate only needs the first 12 bytes and the length -- literals are
// at most 12 bytes long, so a longer value that agrees with one on 12 bytes is decided by its length -- hence: fetch
// bytes 4..11 and go on as if the view were inline.  Rare path: the page is only worked out here -- kStreamItem: `at` is
// the index of the tile in its CTA's stream of the streaming kernel (tiles arrive page by page, pages blockIdx.x,
// blockIdx.x + gridDim.x, ...); else `at` is the page itself.
template <bool kStreamItem = false>
__device__ __forceinline__ uint4 view_first12(uint4 v, const DevPlan& P, uint32_t at) {
  if (__builtin_expect(v.x > 12u, 0)) {
    const uint32_t page = kStreamItem ? blockIdx.x + (at / P.tiles_per_page) * gridDim.x : at;
    const LayoutClass* lc = P.single_class ? &P.class0 : P.classes + P.descs[page].layout_class;
    const uint8_t* p = P.pages + uint64_t(page) * P.page_stride + lc->pool_base + v.w + 4u;
    uint32_t a = 0, b = 0;
#pragma unroll
    for (uint32_t k = 0; k < 4; ++k) {
      a |= uint32_t(p[k]) << (8u * k);
      b |= uint32_t(p[4 + k]) << (8u * k);
    }
    v.z = a;
    v.w = b;
  }
  return v;
}
__device__ __forceinline__ bool view_in_range(uint4 v, const DevTerm& T, const DevPlan& P, uint32_t item) {
  v = view_first12<true>(v, P, item);
  return in_range2((uint64_t(bswap32(v.y)) << 32) | bswap32(v.z), (uint64_t(bswap32(v.w)) << 32) | v.x, T);
}

// Evaluates the conjunct for two rows of the tile with one dispatch (ILP across the rows).
// LD >= 0: the term's load kind is a compile-time constant and the term is a plain range.
template <int LD, bool NONULL>
__device__ __forceinline__ void term_pass2(const DevTerm& T, const uint8_t* stage, uint32_t r0, uint32_t r1,
                                           uint32_t tile_nulls, const DevPlan& P, uint32_t item, bool& p0, bool& p1) {
  const uint8_t* p = stage + T.ref.off;
  bool a, b;
  const uint32_t ld = LD >= 0 ? uint32_t(LD) : uint32_t(T.ref.ld);
  switch (ld) {
    case LD_F64: {
      const int64_t x0 = reinterpret_cast<const int64_t*>(p)[r0], x1 = reinterpret_cast<const int64_t*>(p)[r1];
      a = in_range1(f64_key(x0), T); b = in_range1(f64_key(x1), T);
      break;
    }
    case LD_VIEW: {
      const uint4 v0 = reinterpret_cast<const uint4*>(p)[r0], v1 = reinterpret_cast<const uint4*>(p)[r1];
      a = view_in_range(v0, T, P, item); b = view_in_range(v1, T, P, item);
      break;
    }
    case LD_I32: {
      const int32_t x0 = reinterpret_cast<const int32_t*>(p)[r0], x1 = reinterpret_cast<const int32_t*>(p)[r1];
      a = in_range1(x0, T); b = in_range1(x1, T);
      break;
    }
    case LD_I64: {
      const int64_t x0 = reinterpret_cast<const int64_t*>(p)[r0], x1 = reinterpret_cast<const int64_t*>(p)[r1];
      a = in_range1(x0, T); b = in_range1(x1, T);
      break;
    }
    case LD_I16: {
      const int16_t x0 = reinterpret_cast<const int16_t*>(p)[r0], x1 = reinterpret_cast<const int16_t*>(p)[r1];
      a = in_range1(x0, T); b = in_range1(x1, T);
      break;
    }
    case LD_F32: {
      const int32_t x0 = reinterpret_cast<const int32_t*>(p)[r0], x1 = reinterpret_cast<const int32_t*>(p)[r1];
      a = in_range1(x0 ^ int32_t(uint32_t(x0 >> 31) >> 1), T); b = in_range1(x1 ^ int32_t(uint32_t(x1 >> 31) >> 1), T);
      break;
    }
    case LD_BOOL: {  // bit-packed, LSB first (bitmap.rs:4-29): the key is the bit
      a = in_range1(int64_t((p[r0 >> 3] >> (r0 & 7u)) & 1u), T); b = in_range1(int64_t((p[r1 >> 3] >> (r1 & 7u)) & 1u), T);
      break;
    }
    default: {  // LD_DEC
      const uint4 v0 = reinterpret_cast<const uint4*>(p)[r0], v1 = reinterpret_cast<const uint4*>(p)[r1];
      // signed hi word -> unsigned order by flipping the sign bit (bounds are flipped on the host)
      a = in_range2(((uint64_t(v0.w) << 32) | v0.z) ^ 0x8000000000000000ull, (uint64_t(v0.y) << 32) | v0.x, T);
      b = in_range2(((uint64_t(v1.w) << 32) | v1.z) ^ 0x8000000000000000ull, (uint64_t(v1.y) << 32) | v1.x, T);
      break;
    }
  }
  if (LD < 0 && T.op != TERM_IN_RANGE) { a = T.op == TERM_NOT_IN_RANGE && !a; b = T.op == TERM_NOT_IN_RANGE && !b; }
  if (!NONULL && T.ref.valid_off != kNoValidity && ((tile_nulls >> T.ref.pcol) & 1)) {  // NULL => not TRUE => dropped
    a &= (stage[T.ref.valid_off + (r0 >> 3)] >> (r0 & 7)) & 1;
    b &= (stage[T.ref.valid_off + (r1 >> 3)] >> (r1 & 7)) & 1;
  }
  p0 &= a;
  p1 &= b;
}

// ---- ProjectionExec / aggregate arguments -------------------------------------------------
template <bool NONULL>
__device__ __forceinline__ bool expr_inputs_valid(const DevExpr& e, const Row& row) {
  if ((NONULL || !(row.tile_nulls & e.null_cols)) && !e.has_payload) return true;
  bool ok = true;
#pragma unroll
  for (uint32_t i = 0; i < 3; ++i)
    if (i < e.nfactors) ok &= ref_valid(e.f[i].ref, row);
  return ok;
}

// Float64: one IEEE operation per node, never contracted into an FMA (explicit _rn intrinsics).
template <int FORM>
__device__ __forceinline__ double eval_expr_f64(const DevExpr& e, const Row& row) {
  const uint8_t* st = row.stage;
  const uint32_t r = row.r;
  const uint32_t form = FORM >= 0 ? uint32_t(FORM) : e.form;
  switch (form) {
    case FORM_X:
      return reinterpret_cast<const double*>(st + e.f[0].ref.off)[r];
    case FORM_XY:
      return __dmul_rn(reinterpret_cast<const double*>(st + e.f[0].ref.off)[r], reinterpret_cast<const double*>(st + e.f[1].ref.off)[r]);
    case FORM_X_CMY:
      return __dmul_rn(reinterpret_cast<const double*>(st + e.f[0].ref.off)[r],
                       __dsub_rn(e.f[1].cf, reinterpret_cast<const double*>(st + e.f[1].ref.off)[r]));
    case FORM_X_CMY_CPZ:
      return __dmul_rn(__dmul_rn(reinterpret_cast<const double*>(st + e.f[0].ref.off)[r],
                                 __dsub_rn(e.f[1].cf, reinterpret_cast<const double*>(st + e.f[1].ref.off)[r])),
                       __dadd_rn(e.f[2].cf, reinterpret_cast<const double*>(st + e.f[2].ref.off)[r]));
    default: {
      double v = 1.0;
#pragma unroll
      for (uint32_t i = 0; i < 3; ++i) {
        if (i < e.nfactors) {
          const DevFactor& f = e.f[i];
          double x = load_f64(f.ref, row);
          if (f.kind == PGF_FACTOR_CONST_MINUS_COL) x = __dsub_rn(f.cf, x);
          else if (f.kind == PGF_FACTOR_CONST_PLUS_COL) x = __dadd_rn(f.cf, x);
          v = i == 0 ? x : __dmul_rn(v, x);
        }
      }
      return v;
    }
  }
}

__device__ __forceinline__ uint64_t eval_expr_i64(const DevExpr& e, const Row& row) {
  uint64_t v = 1;
#pragma unroll
  for (uint32_t i = 0; i < 3; ++i) {
    if (i < e.nfactors) {
      const DevFactor& f = e.f[i];
      uint64_t x = uint64_t(load_i64(f.ref, row));
      if (f.kind == PGF_FACTOR_CONST_MINUS_COL) x = uint64_t(f.ci_lo) - x;
      else if (f.kind == PGF_FACTOR_CONST_PLUS_COL) x = uint64_t(f.ci_lo) + x;
      v = i == 0 ? x : v * x;  // wrapping
    }
  }
  return v;
}

__device__ __forceinline__ I128 eval_expr_i128(const DevExpr& e, const Row& row) {
  I128 v{1, 0};
#pragma unroll
  for (uint32_t i = 0; i < 3; ++i) {
    if (i < e.nfactors) {
      const DevFactor& f = e.f[i];
      I128 x;
      if (f.ref.ld == LD_DEC) {
        const uint4 raw = load_u128(f.ref, row);
        x.lo = (uint64_t(raw.y) << 32) | raw.x;
        x.hi = (uint64_t(raw.w) << 32) | raw.z;
      } else {
        const int64_t s = load_i64(f.ref, row);
        x.lo = uint64_t(s);
        x.hi = s < 0 ? ~0ull : 0ull;
      }
      const I128 cst{uint64_t(f.ci_lo), uint64_t(f.ci_hi)};
      if (f.kind == PGF_FACTOR_CONST_MINUS_COL) x = i128_sub(cst, x);
      else if (f.kind == PGF_FACTOR_CONST_PLUS_COL) x = i128_add(cst, x);
      v = i == 0 ? x : i128_mul(v, x);
    }
  }
  return v;
}

template <uint32_t ACC, int FORM>
__device__ __forceinline__ typename AccOps<ACC>::T eval_expr(const DevExpr& e, const Row& row) {
  if constexpr (ACC == CLS_F64) return eval_expr_f64<FORM>(e, row);
  else if constexpr (ACC == CLS_I64) return eval_expr_i64(e, row);
  else return eval_expr_i128(e, row);
}

template <class T>
__device__ __forceinline__ T prev_times_cpz(T prev, const DevExpr& e, const Row& row) {
  if constexpr (std::is_same<T, double>::value)
    return __dmul_rn(prev, __dadd_rn(e.f[2].cf, reinterpret_cast<const double*>(row.stage + e.f[2].ref.off)[row.r]));
  else
    return prev;
}

// ---- global group table ---------------------------------------------------------------
__device__ __forceinline__ uint64_t key_hash(const uint64_t* key, uint32_t nwords, uint32_t knull) {
  uint64_t h = 0x9E3779B97F4A7C15ull ^ knull;
#pragma unroll
  for (uint32_t w = 0; w < kKeyWords; ++w)
    if (w < nwords) h = mix64(h ^ key[w]) + 0x632BE59BD9B4E019ull;
  return h;
}

// Cheap fingerprint for the CTA dictionary: a collision only costs the full compare.
__device__ __forceinline__ uint64_t key_fingerprint(const uint64_t* key, uint32_t knull) {
  uint64_t h = key[0] ^ knull;
  h ^= (key[1] << 17) | (key[1] >> 47);
  h ^= (key[2] << 31) | (key[2] >> 33);
  h ^= (key[3] << 47) | (key[3] >> 17);
  return mix64(h) | 1ull;
}

// The thread that claims a slot writes its key AND zeroes its accumulators and counts before it publishes the slot:
// only the state words of a large table are cleared up front (a 32 M-slot table has 128 MB of state words and
// 1.8 GB of keys / accumulators / counts; clearing all of it cost 0.29 ms per Q3 pass at SF100).
__device__ __forceinline__ void group_slot_init(const GroupTable& t, uint32_t i, const uint64_t* key, uint32_t nwords) {
#pragma unroll
  for (uint32_t w = 0; w < kKeyWords; ++w)
    if (w < nwords) t.keys[uint64_t(i) * kKeyWords + w] = key[w];
  // (fixed trip counts, predicated stores at immediate offsets from one base each: a run-time loop here cost the
  // compaction pipeline 84 bytes of spills)
  const uint32_t nacc = t.nexprs * t.acc_words, ncnt = t.nexprs + 1u;
  uint64_t* acc = t.acc + uint64_t(i) * nacc;
  uint64_t* cnt = t.cnt + uint64_t(i) * ncnt;
#pragma unroll
  for (uint32_t w = 0; w < 2u * kMaxExprs; ++w)
    if (w < nacc) acc[w] = 0ull;
#pragma unroll
  for (uint32_t w = 0; w < kMaxExprs + 1u; ++w)
    if (w < ncnt) cnt[w] = 0ull;
}

// A slot's state word is read with acquire semantics (it pairs with the fence + exchange that publishes the slot):
// the key words read after it are the published ones.  A plain load followed by __threadfence() did the same job
// but made every lookup wait for the thread's own outstanding atomics (4 % of the Q3 lineitem pipeline's stall
// samples sat on those fences).
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Returns the slot holding `key` (inserting it if absent) or -1 when the table is full.
__device__ __forceinline__ int64_t group_slot(const GroupTable& t, const uint64_t* key, uint32_t nwords, uint32_t knull) {
  uint32_t i = uint32_t(key_hash(key, nwords, knull)) & t.mask;
  const uint32_t ready = 2u | (knull << 8);
  for (uint32_t probes = 0; probes <= t.mask; ++probes, i = (i + 1) & t.mask) {
    uint32_t s = ld_acquire_u32(t.state + i);
    if (s == 0) {
      const uint32_t old = atomicCAS(t.state + i, 0u, 1u);
      if (old == 0) {
        group_slot_init(t, i, key, nwords);
        __threadfence();
        atomicExch(t.state + i, ready);
        atomicAdd(t.used, 1u);
        return i;
      }
      s = old == 1u ? old : ld_acquire_u32(t.state + i);   // (the CAS itself is relaxed)
    }
    while ((s & 3u) == 1u) s = ld_acquire_u32(t.state + i);  // being published
    if (s == ready) {
      bool same = true;
#pragma unroll
      for (uint32_t w = 0; w < kKeyWords; ++w)
        if (w < nwords) same &= *reinterpret_cast<volatile uint64_t*>(t.keys + uint64_t(i) * kKeyWords + w) == key[w];
      if (same) return i;
    }
  }
  atomicExch(t.overflow, 1u);
  return -1;
}

// ---- shared-memory block state ---------------------------------------------------------
struct BlockShared {
  uint64_t full[kStages];
  uint64_t empty[kStages];
  StageMeta meta[kStages];
  // CTA-local dictionary of the first kRegGroups group keys (register pre-aggregation)
  uint64_t dict_keys[kRegGroups][kKeyWords];
  uint64_t dict_hash[kRegGroups];
  uint32_t dict_null[kRegGroups];
  uint32_t dict_n;
  uint32_t dict_lock;
  // block reduction scratch
  uint64_t red[kMaxConsumerWarps][2];
  // per-warp deferred-sink queues (entries live behind the stage ring)
  uint32_t qcount[kMaxConsumerWarps];
  uint32_t last_cta;  // this CTA was the last to finish: it runs the fixed-order Float64 reduction
};

__device__ __forceinline__ bool dict_entry_equals(const BlockShared* sh, uint32_t g, const uint64_t* key, uint32_t nwords, uint32_t knull) {
  bool same = *reinterpret_cast<const volatile uint32_t*>(&sh->dict_null[g]) == knull;
#pragma unroll
  for (uint32_t w = 0; w < kKeyWords; ++w)
    if (w < nwords) same &= *reinterpret_cast<const volatile uint64_t*>(&sh->dict_keys[g][w]) == key[w];
  return same;
}

// Find the key in the CTA dictionary, appending it while there is room (rare path: at most
// kRegGroups successful appends per CTA).  Returns -1 when the key is not one of the
// register-resident groups.
struct Key4 {
  uint64_t w0, w1, w2, w3;
};
// Key passed by value so the caller's key words never live in local memory.
static __device__ __noinline__ int dict_lookup_or_insert(BlockShared* sh, Key4 kv, uint32_t nwords, uint32_t knull, uint64_t h) {
  const uint64_t key[kKeyWords] = {kv.w0, kv.w1, kv.w2, kv.w3};
  for (;;) {
    const uint32_t n = *reinterpret_cast<volatile uint32_t*>(&sh->dict_n);
    int found = -1;
#pragma unroll
    for (uint32_t g = 0; g < kRegGroups; ++g)
      if (g < n && dict_entry_equals(sh, g, key, nwords, knull)) found = int(g);
    if (found >= 0) return found;
    if (n == kRegGroups) return -1;
    if (atomicCAS(&sh->dict_lock, 0u, 1u) == 0u) {
      int g = -1;
      if (*reinterpret_cast<volatile uint32_t*>(&sh->dict_n) == n) {  // nothing appended meanwhile
#pragma unroll
        for (uint32_t w = 0; w < kKeyWords; ++w) sh->dict_keys[n][w] = w < nwords ? key[w] : 0;
        sh->dict_null[n] = knull;
        sh->dict_hash[n] = h;
        __threadfence_block();
        *reinterpret_cast<volatile uint32_t*>(&sh->dict_n) = n + 1;
        g = int(n);
      }
      __threadfence_block();
      atomicExch(&sh->dict_lock, 0u);
      if (g >= 0) return g;
    }
  }
}

// ---- register-group fast path (compile-time shapes, NOT NULL scan columns, no join) ------
// Everything below is straight-line code for the two rows a thread handles per iteration, so
// the scheduler can overlap the shared-memory latencies of one row with the arithmetic of the
// other.  The group of a row is resolved against the CTA dictionary with a 32-bit fingerprint
// held in registers plus one exact 128/256-bit compare; the register accumulators are updated
// with predicated adds (no divergent blocks, no jump table).
template <class KEYS>
struct ShapeKeyInfo {
  template <int I>
  static constexpr int words_until() {  // key words used by parts [0, I)
    if constexpr (I == 0) return 0;
    else {
      constexpr int enc = KEYS::template at<I - 1>();
      constexpr int ld = enc & 15;
      constexpr int end = (enc >> 8) + ((ld == LD_VIEW || ld == LD_DEC) ? 2 : 1);
      constexpr int prev = words_until<I - 1>();
      return end > prev ? end : prev;
    }
  }
  static constexpr int nwords = words_until<KEYS::size>();
};

template <class KEYS>
__device__ __forceinline__ void load_key_fast(const DevPlan& P, const uint8_t* stage, uint32_t r, uint64_t (&key)[kKeyWords],
                                              uint32_t& bad) {
  static_for<KEYS::size>([&](auto I) {
    constexpr int kp = decltype(I)::value;
    constexpr int enc = KEYS::template at<kp>();
    constexpr int ld = enc & 15;
    constexpr int word = enc >> 8;
    const uint8_t* col = stage + P.keys[kp].ref.off;
    if constexpr (ld == LD_VIEW || ld == LD_DEC) {
      const uint4 raw = reinterpret_cast<const uint4*>(col)[r];
      if (ld == LD_VIEW) bad += raw.x > 12u;
      key[word] = (uint64_t(raw.y) << 32) | raw.x;
      key[word + 1 < int(kKeyWords) ? word + 1 : word] = (uint64_t(raw.w) << 32) | raw.z;
    } else if constexpr (ld == LD_I64) {
      key[word] = uint64_t(reinterpret_cast<const int64_t*>(col)[r]);
    } else if constexpr (ld == LD_I32) {
      key[word] = uint64_t(int64_t(reinterpret_cast<const int32_t*>(col)[r]));
    } else {
      key[word] = uint64_t(int64_t(reinterpret_cast<const int16_t*>(col)[r]));
    }
  });
}

// 32-bit fingerprint over the used key words; odd, so it never equals an empty (0) entry.
template <int NW>
__device__ __forceinline__ uint32_t key_fp32(const uint64_t (&key)[kKeyWords]) {
  uint32_t f = 0x9E3779B9u;
#pragma unroll
  for (int w = 0; w < NW; ++w) {
    const uint32_t lo = uint32_t(key[w]), hi = uint32_t(key[w] >> 32);
    f = __funnelshift_l(f, f, 5) ^ lo ^ __funnelshift_l(hi, hi, 13);
  }
  return f | 1u;
}

template <int NW>
__device__ __forceinline__ bool dict_equal_fast(const BlockShared* sh, int g, const uint64_t (&key)[kKeyWords]) {
  const ulonglong2* e = reinterpret_cast<const ulonglong2*>(&sh->dict_keys[g < 0 ? 0 : g][0]);
  uint64_t diff = 0;
  const ulonglong2 a = e[0];
  diff |= a.x ^ key[0];
  if (NW > 1) diff |= a.y ^ key[1];
  if constexpr (NW > 2) {
    const ulonglong2 b = e[1];
    diff |= b.x ^ key[2];
    if (NW > 3) diff |= b.y ^ key[3];
  }
  return g >= 0 && diff == 0;
}

// Float64 argument forms over NOT NULL scan columns (compile-time FORM)
template <int FORM>
__device__ __forceinline__ double eval_fast_f64(const DevExpr& e, const uint8_t* stage, uint32_t r, double prev) {
  auto col = [&](int f) { return reinterpret_cast<const double*>(stage + e.f[f].ref.off)[r]; };
  if constexpr (FORM == int(FORM_X)) return col(0);
  else if constexpr (FORM == int(FORM_XY)) return __dmul_rn(col(0), col(1));
  else if constexpr (FORM == int(FORM_X_CMY)) return __dmul_rn(col(0), __dsub_rn(e.f[1].cf, col(1)));
  else if constexpr (FORM == int(FORM_X_CMY_CPZ)) return __dmul_rn(__dmul_rn(col(0), __dsub_rn(e.f[1].cf, col(1))), __dadd_rn(e.f[2].cf, col(2)));
  else return __dmul_rn(prev, __dadd_rn(e.f[2].cf, col(2)));  // FORM_PREV_CPZ
}

// Decimal128 argument forms over NOT NULL Decimal128 scan columns: wrapping i128, no rescale
template <int FORM>
__device__ __forceinline__ I128 eval_fast_i128(const DevExpr& e, const uint8_t* stage, uint32_t r, I128 prev) {
  auto col = [&](int f) {
    const uint4 raw = reinterpret_cast<const uint4*>(stage + e.f[f].ref.off)[r];
    return I128{(uint64_t(raw.y) << 32) | raw.x, (uint64_t(raw.w) << 32) | raw.z};
  };
  auto cst = [&](int f) { return I128{uint64_t(e.f[f].ci_lo), uint64_t(e.f[f].ci_hi)}; };
  if constexpr (FORM == int(FORM_X)) return col(0);
  else if constexpr (FORM == int(FORM_XY)) return i128_mul_narrow(col(0), col(1));
  else if constexpr (FORM == int(FORM_X_CMY)) return i128_mul_narrow(col(0), i128_sub(cst(1), col(1)));
  else if constexpr (FORM == int(FORM_X_CMY_CPZ)) return i128_mul_narrow(i128_mul_narrow(col(0), i128_sub(cst(1), col(1))), i128_add(cst(2), col(2)));
  else return i128_mul_narrow(prev, i128_add(cst(2), col(2)));  // FORM_PREV_CPZ
}

template <uint32_t ACC, int FORM>
__device__ __forceinline__ typename AccOps<ACC>::T eval_fast(const DevExpr& e, const uint8_t* stage, uint32_t r, typename AccOps<ACC>::T prev) {
  if constexpr (ACC == CLS_F64) return eval_fast_f64<FORM>(e, stage, r, prev);
  else return eval_fast_i128<FORM>(e, stage, r, prev);
}

// One 8-byte shared-memory accumulator slot.  Float64: the running sum.  Decimal128: a signed 64-bit
// partial sum -- values that do not fit 64 bits, or an addition that would overflow the slot, are
// reported (false) and go to the 128-bit accumulator of the global table instead, so the total is the
// exact wrapping i128 sum whatever the magnitudes.
template <uint32_t ACC>
__device__ __forceinline__ bool slot_add(unsigned long long* slot, typename AccOps<ACC>::T v) {
  if constexpr (ACC == CLS_F64) {
    *reinterpret_cast<double*>(slot) = __dadd_rn(*reinterpret_cast<double*>(slot), v);
    return true;
  } else {
    const int64_t x = int64_t(v.lo), s = int64_t(*slot);
    const int64_t r = int64_t(uint64_t(s) + uint64_t(x));
    const bool fits = v.hi == uint64_t(x >> 63);
    const bool overflow = ((s ^ r) & (x ^ r)) < 0;
    if (fits && !overflow) *slot = uint64_t(r);
    return fits && !overflow;
  }
}
template <uint32_t ACC>
__device__ __forceinline__ typename AccOps<ACC>::T slot_value(const unsigned long long* slot) {
  if constexpr (ACC == CLS_F64) return *reinterpret_cast<const double*>(slot);
  else {
    const int64_t s = int64_t(*slot);
    return I128{uint64_t(s), s < 0 ? ~0ull : 0ull};
  }
}

// Rare paths of the fast sink.  (a) The row's group is not slot resident (more than kRegGroups groups
// in this CTA): the whole row goes straight to the global table.  (b) One Decimal128 value could not
// be added to its 64-bit slot: only that value goes to the global accumulator (the row is still
// counted by its slot).
template <uint32_t ACC, int NE>
struct FastValues {
  typename AccOps<ACC>::T v[NE];
};
template <uint32_t ACC, int NE>
static __device__ __noinline__ void fast_slow_accumulate(const DevPlan& P, Key4 kv, FastValues<ACC, NE> vals) {
  using Ops = AccOps<ACC>;
  const uint64_t key[kKeyWords] = {kv.w0, kv.w1, kv.w2, kv.w3};
  const int64_t slot = group_slot(P.table, key, P.nkeywords, 0);
  if (slot < 0) return;
#pragma unroll
  for (int e = 0; e < NE; ++e) {
    Ops::atomic_add(P.table.acc + (uint64_t(slot) * P.nexprs + e) * P.table.acc_words, vals.v[e]);
    atomicAdd(reinterpret_cast<unsigned long long*>(P.table.cnt + uint64_t(slot) * (P.nexprs + 1) + e), 1ull);
  }
  atomicAdd(reinterpret_cast<unsigned long long*>(P.table.cnt + uint64_t(slot) * (P.nexprs + 1) + P.nexprs), 1ull);
}
template <uint32_t ACC>
static __device__ __noinline__ void fast_value_to_global(const DevPlan& P, Key4 kv, uint32_t e, typename AccOps<ACC>::T v) {
  const uint64_t key[kKeyWords] = {kv.w0, kv.w1, kv.w2, kv.w3};
  const int64_t slot = group_slot(P.table, key, P.nkeywords, 0);
  if (slot >= 0) AccOps<ACC>::atomic_add(P.table.acc + (uint64_t(slot) * P.nexprs + e) * P.table.acc_words, v);
}

// One row straight into the global group table (NULL inputs, groups beyond the register set).
template <uint32_t ACC, uint32_t MAXE>
__device__ __forceinline__ void global_accumulate(const DevPlan& P, bool grouped, const uint64_t* key, uint32_t knull,
                                                  uint32_t valid_mask, const typename AccOps<ACC>::T* v) {
  using Ops = AccOps<ACC>;
  const int64_t slot = grouped ? group_slot(P.table, key, P.nkeywords, knull) : 0;
  if (slot < 0) return;
#pragma unroll
  for (uint32_t e = 0; e < MAXE; ++e) {
    if (e < P.nexprs && ((valid_mask >> e) & 1)) {
      Ops::atomic_add(P.table.acc + (uint64_t(slot) * P.nexprs + e) * P.table.acc_words, v[e]);
      atomicAdd(reinterpret_cast<unsigned long long*>(P.table.cnt + uint64_t(slot) * (P.nexprs + 1) + e), 1ull);
    }
  }
  atomicAdd(reinterpret_cast<unsigned long long*>(P.table.cnt + uint64_t(slot) * (P.nexprs + 1) + P.nexprs), 1ull);
}

// ---- the kernel ----------------------------------------------------------------------
// straight-line two-row sink: registered shape over NOT NULL scan columns, Float64 or Decimal128 sums of
// arguments with a compile-time form, no join
template <uint32_t SINK, uint32_t ACC, bool GROUPED, uint32_t NJ, class SHAPE>
constexpr bool is_fast_grouped() {
  return SINK == SINK_AGG && GROUPED && NJ == 0 && (ACC == CLS_F64 || ACC == CLS_I128) && !SHAPE::generic && SHAPE::no_nulls &&
         SHAPE::Keys::size > 0 && SHAPE::Exprs::size > 0 && SHAPE::Exprs::template at<0>() != int(FORM_GENERIC);
}

template <uint32_t SINK, uint32_t ACC, bool GROUPED, uint32_t NJ, uint32_t MAXE_T, class SHAPE = GenericShape>
__global__ void __launch_bounds__(pipeline_threads(SINK, GROUPED, is_fast_grouped<SINK, ACC, GROUPED, NJ, SHAPE>(), ACC), 1)
pipeline_kernel(const __grid_constant__ DevPlan P) {
  constexpr bool kFastGrouped = is_fast_grouped<SINK, ACC, GROUPED, NJ, SHAPE>();
  constexpr int kConsumerWarps = consumer_warps(SINK, GROUPED, kFastGrouped, ACC);
  constexpr uint32_t PW = uint32_t(producer_warps(kFastGrouped, ACC));   // warps 0 .. PW-1 produce, the rest consume
  using Ops = AccOps<ACC>;
  using AccT = typename Ops::T;
  constexpr uint32_t MAXE = SINK == SINK_AGG ? MAXE_T : 1;
  constexpr uint32_t G = (SINK == SINK_AGG && GROUPED) ? kRegGroups : 1;
  constexpr uint32_t kRows = rows_per_thread(NJ);  // rows per thread and iteration
  constexpr uint32_t kAccThreads = uint32_t(kConsumerWarps) * 32u;  // threads with shared-memory accumulator slots
  const uint32_t kNumStages = P.nstages;  // 3 or kStages
  constexpr uint32_t kAccWords = ACC == CLS_I128 ? 2 : 1;
  [[maybe_unused]] constexpr uint32_t kQueueEntryWords = queue_entry_words(MAXE, kAccWords);
  [[maybe_unused]] constexpr uint32_t kQueueCap = kQueueBytesPerWarp / (8 * kQueueEntryWords);
  [[maybe_unused]] constexpr uint32_t kQueueDrainAt = kQueueCap > 12 ? kQueueCap - 8 : kQueueCap / 2;

  extern __shared__ __align__(128) uint8_t smem_raw[];
  BlockShared* sh = reinterpret_cast<BlockShared*>(smem_raw);
  uint8_t* stages = smem_raw + ((sizeof(BlockShared) + 127) & ~size_t(127));

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&sh->full[s], 1);
      mbar_init(&sh->empty[s], kConsumerWarps);
    }
    sh->dict_n = 0;
    sh->dict_lock = 0;
    for (int w = 0; w < kMaxConsumerWarps; ++w) sh->qcount[w] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  // per-thread state that outlives the tile loop
  AccT acc[G][MAXE];
  uint32_t grows[G];
#pragma unroll
  for (uint32_t g = 0; g < G; ++g) {
    grows[g] = 0;
#pragma unroll
    for (uint32_t e = 0; e < MAXE; ++e) acc[g][e] = Ops::zero();
  }
  uint64_t dh[G];  // hashes of the CTA dictionary entries, cached in registers
  uint32_t dn = 0;
#pragma unroll
  for (uint32_t g = 0; g < G; ++g) dh[g] = 0;
  uint32_t n_in = 0, n_bloom = 0, n_filt = 0, n_out = 0, n_bad = 0, n_bloom_ins = 0;

  // shared-memory accumulator slots of the fast GROUP BY path (behind the ring and the queues)
  [[maybe_unused]] unsigned long long* myacc =
      reinterpret_cast<unsigned long long*>(stages + size_t(kNumStages) * P.stage_bytes) +   // (the fast path has no deferred-sink queues)
      (threadIdx.x >= PW * 32u ? threadIdx.x - PW * 32u : 0);
  if constexpr (kFastGrouped) {
    if (warp >= PW)
      for (uint32_t q = 0; q < G * (SHAPE::Exprs::size + 1); ++q) myacc[q * kAccThreads] = 0ull;  // +0.0 / integer 0
  }

  if (warp < PW) {
    // ===== producer(s): TMA bulk copies of the needed column slices of each row tile =====
    // One warp feeds the whole CTA, so its per-tile path is kept short: pages are walked with a
    // nested page / tile loop (no divisions), each lane keeps its column's plan entry and the
    // offsets of the current layout class in registers, the byte count is one warp reduction.
    uint32_t ps = 0, pphase = 0;  // ring stage and mbarrier phase of the producer
    [[maybe_unused]] uint32_t tseq = 0;  // tile sequence number inside the CTA
    // lanes 0..15: values slice of staged column `lane`; lanes 16..31: its validity slice
    const uint32_t mycol = lane & 15u;
    const bool is_validity = lane >= 16;
    const bool has_col = mycol < P.nstage_cols;
    DevStageCol sc{};
    if (has_col) sc = P.scol[mycol];
    const bool want = has_col && (!is_validity || sc.nullable);
    const uint32_t width = is_validity ? 0u : uint32_t(sc.width);
    const uint32_t smem_off = is_validity ? sc.valid_off : sc.smem_off;
    uint32_t cur_class = 0xFFFFFFFFu, col_off = 0;
    uint32_t page = blockIdx.x;
    PageDesc d_next{};
    if (page < P.npages) d_next = P.descs[page];
    for (; page < P.npages; page += gridDim.x) {
      const PageDesc d = d_next;
      if (page + gridDim.x < P.npages) d_next = P.descs[page + gridDim.x];  // prefetch
      if (d.layout_class != cur_class) {  // rare: pages of one scan share their layout class
        cur_class = d.layout_class;
        const LayoutClass* lc = P.classes + cur_class;
        col_off = want ? (is_validity ? lc->validity_off[sc.page_col] : lc->values_off[sc.page_col]) : 0u;
      }
      const uint8_t* col_base = P.pages + uint64_t(page) * P.page_stride + col_off;
      const uint32_t null_mask = d.null_mask & P.used_null_mask;
      const bool active = want && (!is_validity || ((null_mask >> sc.page_col) & 1u));
      for (uint32_t tile = 0, r0 = 0; tile < P.tiles_per_page; ++tile, r0 += P.tile_rows) {
        const uint32_t s = ps;
        const uint32_t wait_parity = pphase ^ 1u;
        if (++ps == kNumStages) { ps = 0; pphase ^= 1u; }
        if (PW > 1 && ((tseq++) % PW) != warp) continue;   // the producers take the CTA's tiles in turn
        const uint32_t n = d.row_count > r0 ? min(d.row_count - r0, P.tile_rows) : 0u;
        uint32_t bytes = 0;
        if (active && n) bytes = width == 0u ? ((((n + 7u) >> 3) + 15u) & ~15u) : ((n * width + 15u) & ~15u);   // (width 0: a bitmap -- validity, or Boolean values)
        const uint32_t total = __reduce_add_sync(0xffffffffu, bytes);
        mbar_wait(&sh->empty[s], wait_parity);
        if (lane == 0) {
          sh->meta[s].nrows = n;
          sh->meta[s].null_mask = null_mask;
          sh->meta[s].row_base = d.row_base + r0;
          mbar_arrive_expect_tx(&sh->full[s], total);
        }
        __syncwarp();
        if (bytes)
          tma_load_1d(stages + size_t(s) * P.stage_bytes + smem_off, col_base + (width == 0u ? (r0 >> 3) : r0 * width), bytes, &sh->full[s]);
      }
    }
  } else {
    // ===== consumers =====
    [[maybe_unused]] uint64_t* myqueue = reinterpret_cast<uint64_t*>(stages + size_t(kNumStages) * P.stage_bytes) +
                                         size_t(warp - PW) * (kQueueBytesPerWarp / 8);
    [[maybe_unused]] auto drain_queue = [&]() {
      if constexpr (SINK == SINK_AGG && GROUPED) {
        __syncwarp();
        const uint32_t n = min(*reinterpret_cast<volatile uint32_t*>(&sh->qcount[warp - PW]), kQueueCap);
        for (uint32_t base = 0; base < n; base += 32u) {
          const uint32_t idx = base + lane;
          if (idx < n) {
            const uint64_t* qe = myqueue + idx * kQueueEntryWords;
            uint64_t key[kKeyWords];
            AccT v[MAXE];
#pragma unroll
            for (uint32_t w = 0; w < kKeyWords; ++w) key[w] = qe[w];
#pragma unroll
            for (uint32_t e = 0; e < MAXE; ++e) v[e] = *reinterpret_cast<const AccT*>(qe + kKeyWords + 1 + e * kAccWords);
            global_accumulate<ACC, MAXE>(P, true, key, uint32_t(qe[kKeyWords]), uint32_t(qe[kKeyWords] >> 32), v);
          }
        }
        __syncwarp();
        if (lane == 0) sh->qcount[warp - PW] = 0;
        __syncwarp();
      }
    };
    uint32_t k = 0, cs = 0, cphase = 0;  // ring stage and mbarrier phase of this consumer
    // tiles arrive in the producer's order: the pages blockIdx.x, blockIdx.x + gridDim.x, ... tile by tile
    const uint32_t my_pages = P.npages > blockIdx.x ? (P.npages - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
    const uint32_t my_items = my_pages * P.tiles_per_page;
    for (uint32_t item = 0; item < my_items; ++item, ++k) {
      const uint32_t s = cs;
      mbar_wait(&sh->full[s], cphase);
      if (++cs == kNumStages) { cs = 0; cphase ^= 1u; }
      const uint32_t nrows = sh->meta[s].nrows;
      const uint8_t* stage = stages + size_t(s) * P.stage_bytes;
      const uint32_t tile_nulls = sh->meta[s].null_mask;
      // R rows per thread and iteration (independent chains => ILP / memory-level parallelism):
      // a warp takes R consecutive 32-row chunks.  R = 2 for streaming pipelines; R = 4 behind a
      // join probe, where every row carries a chain of dependent L2/HBM round trips and the rows
      // in flight per SM bound the throughput.  Chunk groups are dealt to the warps round-robin
      // with a per-tile rotation so partial tiles do not always load the same warps.  The loop
      // bound is warp-uniform: the warp votes below are executed by all 32 lanes.
      constexpr uint32_t R = kRows;
      const uint32_t ngroups = (nrows + 32u * R - 1u) / (32u * R);
      for (uint32_t pr = (warp - PW + kConsumerWarps - (k * 5u) % kConsumerWarps) % kConsumerWarps; pr < ngroups; pr += kConsumerWarps) {
        const uint32_t b0 = pr * (32u * R);
        uint32_t rr[R];
        bool keep[R];
#pragma unroll
        for (uint32_t q = 0; q < R; ++q) {
          const bool has = b0 + 32u * q + lane < nrows;
          rr[q] = has ? b0 + 32u * q + lane : 0u;  // row 0 of a tile always exists
          keep[q] = has;
          n_in += uint32_t(has);
        }
        // -- runtime Bloom probes: NULL key => DefinitelyAbsent (shared.rs:367-374)
        for (uint32_t b = 0; b < P.nbloom; ++b) {
          const DevBloomProbe& bp = P.bloom[b];
          uint64_t bk[R];
#pragma unroll
          for (uint32_t q = 0; q < R; ++q) {
            const Row rq{stage, rr[q], tile_nulls, nullptr, 0};
            keep[q] = keep[q] && ref_valid(bp.key, rq);
            bk[q] = uint64_t(load_i64(bp.key, rq));
          }
          bloom_contains_n<R>(bp.bloom, bk, keep);
        }
        bool any_keep = false;
#pragma unroll
        for (uint32_t q = 0; q < R; ++q) { n_bloom += uint32_t(keep[q]); any_keep |= keep[q]; }
        // -- FilterExec: every conjunct must be TRUE; stop as soon as the whole warp is dead
        if constexpr (SHAPE::generic) {
          for (uint32_t t = 0; t < P.nterms; ++t) {
            if (!__any_sync(0xffffffffu, any_keep)) break;
            any_keep = false;
#pragma unroll
            for (uint32_t q = 0; q < R; q += 2) {
              term_pass2<-1, false>(P.terms[t], stage, rr[q], rr[q + 1], tile_nulls, P, item, keep[q], keep[q + 1]);
              any_keep |= keep[q] || keep[q + 1];
            }
          }
        } else {
          static_for<SHAPE::Terms::size>([&](auto I) {
            constexpr int t = decltype(I)::value;
            if (__any_sync(0xffffffffu, any_keep)) {
              any_keep = false;
#pragma unroll
              for (uint32_t q = 0; q < R; q += 2) {
                term_pass2<SHAPE::Terms::template at<t>(), SHAPE::no_nulls>(P.terms[t], stage, rr[q], rr[q + 1], tile_nulls, P, item, keep[q], keep[q + 1]);
                any_keep |= keep[q] || keep[q + 1];
              }
            }
          });
        }
#pragma unroll
        for (uint32_t q = 0; q < R; ++q) n_filt += uint32_t(keep[q]);

        if (!__any_sync(0xffffffffu, any_keep)) continue;
        [[maybe_unused]] const uint32_t r0 = rr[0], r1 = rr[1];
        [[maybe_unused]] const bool keep0 = keep[0], keep1 = keep[1];
        if constexpr (kFastGrouped) {
          constexpr int NW = ShapeKeyInfo<typename SHAPE::Keys>::nwords;
          constexpr int NE = SHAPE::Exprs::size;
          n_out += uint32_t(keep0) + uint32_t(keep1);
          uint64_t key0[kKeyWords] = {0, 0, 0, 0}, key1[kKeyWords] = {0, 0, 0, 0};
          load_key_fast<typename SHAPE::Keys>(P, stage, r0, key0, n_bad);
          load_key_fast<typename SHAPE::Keys>(P, stage, r1, key1, n_bad);
          const uint32_t f0 = key_fp32<NW>(key0), f1 = key_fp32<NW>(key1);
          int g0 = -1, g1 = -1;
#pragma unroll
          for (uint32_t gg = 0; gg < G; ++gg) {
            if (uint32_t(dh[gg]) == f0) g0 = int(gg);
            if (uint32_t(dh[gg]) == f1) g1 = int(gg);
          }
          const bool hit0 = dict_equal_fast<NW>(sh, g0, key0), hit1 = dict_equal_fast<NW>(sh, g1, key1);
          if ((keep0 && !hit0) || (keep1 && !hit1)) {  // first rows of a group in this CTA, or > kRegGroups groups
            if (keep0 && !hit0) g0 = dict_lookup_or_insert(sh, Key4{key0[0], key0[1], key0[2], key0[3]}, uint32_t(NW), 0u, f0);
            if (keep1 && !hit1) g1 = dict_lookup_or_insert(sh, Key4{key1[0], key1[1], key1[2], key1[3]}, uint32_t(NW), 0u, f1);
            const uint32_t n = *reinterpret_cast<volatile uint32_t*>(&sh->dict_n);
#pragma unroll
            for (uint32_t gg = 0; gg < G; ++gg) dh[gg] = gg < n ? *reinterpret_cast<volatile uint64_t*>(&sh->dict_hash[gg]) : 0;
          }
          AccT v0[NE], v1[NE];
          static_for<NE>([&](auto I) {
            constexpr int e = decltype(I)::value;
            constexpr int form = SHAPE::Exprs::template at<e>();
            v0[e] = eval_fast<ACC, form>(P.exprs[e], stage, r0, v0[e > 0 ? e - 1 : 0]);
            v1[e] = eval_fast<ACC, form>(P.exprs[e], stage, r1, v1[e > 0 ? e - 1 : 0]);
          });
          const int a0 = keep0 ? g0 : -2, a1 = keep1 ? g1 : -2;
          // Accumulators of the slot-resident groups live in shared memory, one private 8-byte slot per
          // (group, argument, thread): `accs[(g * (NE + 1) + e) * T + tid]`.  Indexing by the row's
          // group costs one load-add-store per argument and no divergence; selecting among register
          // accumulators costs three instructions per (group, argument) -- 64 vs 20 per row for Q1.
          uint32_t spill0 = 0, spill1 = 0;  // Decimal128 values that did not fit their slot
          if (a0 >= 0) {
            unsigned long long* p0 = myacc + uint32_t(a0) * ((NE + 1) * kAccThreads);
#pragma unroll
            for (int e = 0; e < NE; ++e) spill0 |= uint32_t(!slot_add<ACC>(p0 + e * kAccThreads, v0[e])) << e;
            p0[NE * kAccThreads] += 1ull;
          }
          if (a1 >= 0) {
            unsigned long long* p1 = myacc + uint32_t(a1) * ((NE + 1) * kAccThreads);
#pragma unroll
            for (int e = 0; e < NE; ++e) spill1 |= uint32_t(!slot_add<ACC>(p1 + e * kAccThreads, v1[e])) << e;
            p1[NE * kAccThreads] += 1ull;
          }
          if (a0 == -1 || a1 == -1) {
            FastValues<ACC, NE> x0, x1;
#pragma unroll
            for (int e = 0; e < NE; ++e) { x0.v[e] = v0[e]; x1.v[e] = v1[e]; }
            if (a0 == -1) fast_slow_accumulate<ACC, NE>(P, Key4{key0[0], key0[1], key0[2], key0[3]}, x0);
            if (a1 == -1) fast_slow_accumulate<ACC, NE>(P, Key4{key1[0], key1[1], key1[2], key1[3]}, x1);
          }
          if constexpr (ACC != CLS_F64) {
            if (spill0 | spill1) {
#pragma unroll
              for (int e = 0; e < NE; ++e) {
                if ((spill0 >> e) & 1) fast_value_to_global<ACC>(P, Key4{key0[0], key0[1], key0[2], key0[3]}, uint32_t(e), v0[e]);
                if ((spill1 >> e) & 1) fast_value_to_global<ACC>(P, Key4{key1[0], key1[1], key1[2], key1[3]}, uint32_t(e), v1[e]);
              }
            }
          }
          continue;
        }
        // (plans with a HashJoinExec probe or a build sink run the compaction pipeline, probe_kernel.cuh)
        static_assert(NJ == 0 && SINK != SINK_JOIN_BUILD, "the streaming kernel has no join stages");
        uint32_t pending = 0;  // bit h: row h reaches the sink
#pragma unroll
        for (uint32_t h = 0; h < R; ++h) pending |= uint32_t(keep[h]) << h;
        while (pending) {
          const uint32_t half = 31u - __clz(pending & (0u - pending));
          pending &= pending - 1u;
          uint32_t rsel = rr[0];
#pragma unroll
          for (uint32_t h = 1; h < R; ++h)
            if (half == h) rsel = rr[h];
          Row row{stage, rsel, tile_nulls, nullptr, 0};

          // -- sink (optionally behind one HashJoinExec probe)
          auto sink = [&](const Row& rc) {
            ++n_out;
            if constexpr (SINK == SINK_AGG) {
              uint64_t key[kKeyWords] = {0, 0, 0, 0};
              uint32_t knull = 0;
              int g = 0;
              if constexpr (GROUPED) {
                if constexpr (!SHAPE::generic && SHAPE::Keys::size > 0) {
                  // key layout known at compile time: every word lands in a fixed register
                  static_for<SHAPE::Keys::size>([&](auto I) {
                    constexpr int kp = decltype(I)::value;
                    constexpr int enc = SHAPE::Keys::template at<kp>();
                    constexpr int ld = enc & 15;
                    constexpr bool pay = ((enc >> 4) & 1) != 0;
                    constexpr int word = enc >> 8;
                    const DevKeyPart& part = P.keys[kp];
                    bool valid = true;
                    if (pay || !SHAPE::no_nulls) valid = ref_valid(part.ref, rc);
                    if (!valid) {
                      knull |= 1u << kp;
                    } else if constexpr (ld == LD_VIEW || ld == LD_DEC) {
                      uint4 raw;
                      if constexpr (pay) {
                        const uint32_t* pp = rc.pay + 3 + part.ref.off;
                        raw = make_uint4(__ldg(pp), __ldg(pp + 1), __ldg(pp + 2), __ldg(pp + 3));
                      } else {
                        raw = reinterpret_cast<const uint4*>(rc.stage + part.ref.off)[rc.r];
                      }
                      if (ld == LD_VIEW && raw.x > 12u) ++n_bad;
                      key[word] = (uint64_t(raw.y) << 32) | raw.x;
                      key[word + 1 < int(kKeyWords) ? word + 1 : word] = (uint64_t(raw.w) << 32) | raw.z;
                    } else {
                      int64_t x;
                      if constexpr (pay) {
                        const uint32_t* pp = rc.pay + 3 + part.ref.off;
                        x = ld == LD_I64 ? int64_t((uint64_t(__ldg(pp + 1)) << 32) | __ldg(pp))
                          : ld == LD_I32 ? int64_t(int32_t(__ldg(pp))) : int64_t(int16_t(__ldg(pp)));
                      } else {
                        const uint8_t* pp = rc.stage + part.ref.off;
                        x = ld == LD_I64 ? reinterpret_cast<const int64_t*>(pp)[rc.r]
                          : ld == LD_I32 ? int64_t(reinterpret_cast<const int32_t*>(pp)[rc.r])
                                         : int64_t(reinterpret_cast<const int16_t*>(pp)[rc.r]);
                      }
                      key[word] = uint64_t(x);
                    }
                  });
                } else {
#pragma unroll
                for (uint32_t kp = 0; kp < 4; ++kp) {
                  if (kp < P.nkeys) {
                    const DevKeyPart& part = P.keys[kp];
                    if (!ref_valid(part.ref, rc)) {
                      knull |= 1u << kp;  // NULL keys form one group; words stay zero
                    } else {
                      uint64_t w0, w1 = 0;
                      if (part.nwords == 1) {
                        w0 = uint64_t(load_i64(part.ref, rc));
                      } else {
                        const uint4 raw = load_u128(part.ref, rc);
                        if (part.ref.ld == LD_VIEW && raw.x > 12u) ++n_bad;
                        w0 = (uint64_t(raw.y) << 32) | raw.x;
                        w1 = (uint64_t(raw.w) << 32) | raw.z;
                      }
#pragma unroll
                      for (uint32_t w = 0; w < kKeyWords; ++w) {
                        if (w == part.word) key[w] = w0;
                        if (part.nwords == 2 && w == part.word + 1u) key[w] = w1;
                      }
                    }
                  }
                }
                }
                // CTA dictionary lookup: hashes cached in registers, full compare on a hit
                const uint64_t h = key_fingerprint(key, knull);
                g = -1;
#pragma unroll
                for (uint32_t gg = 0; gg < G; ++gg)
                  if (gg < dn && dh[gg] == h) g = int(gg);
                // the dictionary is append-only: once it is full a key without a fingerprint hit is not in it
                if (g < 0 && dn == G) {
                } else if (g < 0 || !dict_entry_equals(sh, uint32_t(g), key, P.nkeywords, knull)) {
                  g = dict_lookup_or_insert(sh, Key4{key[0], key[1], key[2], key[3]}, P.nkeywords, knull, h);
                  dn = *reinterpret_cast<volatile uint32_t*>(&sh->dict_n);
#pragma unroll
                  for (uint32_t gg = 0; gg < G; ++gg) dh[gg] = gg < dn ? *reinterpret_cast<volatile uint64_t*>(&sh->dict_hash[gg]) : 0;
                }
              }
              bool all_valid = true;
              uint32_t valid_mask = 0;
              AccT v[MAXE];
              static_for<int(MAXE)>([&](auto I) {
                constexpr int e = decltype(I)::value;
                constexpr int form = SHAPE::generic ? -1 : SHAPE::Exprs::template at<e>();  // FORM_PREV_CPZ only in shapes
                v[e] = Ops::zero();
                if (SHAPE::generic ? uint32_t(e) < P.nexprs : e < SHAPE::Exprs::size) {
                  const bool ok = expr_inputs_valid<SHAPE::no_nulls>(P.exprs[e], rc);
                  if constexpr (ACC == CLS_F64 && form == int(FORM_PREV_CPZ) && e > 0) {
                    // common subexpression: this argument is the previous one times (c + z)
                    v[e] = prev_times_cpz(v[e > 0 ? e - 1 : 0], P.exprs[e], rc);
                  } else {
                    v[e] = eval_expr<ACC, form>(P.exprs[e], rc);
                  }
                  all_valid &= ok;
                  valid_mask |= uint32_t(ok) << e;
                }
              });
              if (all_valid && g >= 0) {
                // fast path: register accumulators of the row's group (one short divergent
                // block per register group present in the warp)
#pragma unroll
                for (uint32_t gg = 0; gg < G; ++gg) {
                  if (G == 1 || gg == uint32_t(g)) {
                    grows[gg] += 1;
                    static_for<int(MAXE)>([&](auto I) {
                      constexpr int e = decltype(I)::value;
                      if (SHAPE::generic ? uint32_t(e) < P.nexprs : e < SHAPE::Exprs::size) acc[gg][e] = Ops::add(acc[gg][e], v[e]);
                    });
                  }
                }
              } else {
                // slow path (NULL inputs, or > kRegGroups groups): queue the evaluated row; the warp
                // applies its queue to the global table from a converged point
                bool queued = false;
                if constexpr (GROUPED) {
                  const uint32_t pos = atomicAdd(&sh->qcount[warp - PW], 1u);
                  if (pos < kQueueCap) {
                    uint64_t* qe = myqueue + pos * kQueueEntryWords;
#pragma unroll
                    for (uint32_t w = 0; w < kKeyWords; ++w) qe[w] = key[w];
                    qe[kKeyWords] = uint64_t(knull) | (uint64_t(valid_mask) << 32);
#pragma unroll
                    for (uint32_t e = 0; e < MAXE; ++e) *reinterpret_cast<AccT*>(qe + kKeyWords + 1 + e * kAccWords) = v[e];
                    queued = true;
                  }
                }
                if (!queued) global_accumulate<ACC, MAXE>(P, GROUPED, key, knull, valid_mask, v);
              }
            }
          };

          sink(row);
        }
        if constexpr (SINK == SINK_AGG && GROUPED) {
          __syncwarp();
          if (*reinterpret_cast<volatile uint32_t*>(&sh->qcount[warp - PW]) >= kQueueDrainAt) drain_queue();
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->empty[s]);
    }
    drain_queue();
  }
  __syncthreads();

  // ===== epilogue: block-reduce the register accumulators, then one atomic per CTA =====
  if constexpr (SINK == SINK_AGG) {
    const uint32_t ng = GROUPED ? sh->dict_n : 1u;
    for (uint32_t g = 0; g < ng; ++g) {
      int64_t slot = 0;
      if (GROUPED && threadIdx.x == PW * 32u) {
        uint64_t key[kKeyWords];
#pragma unroll
        for (uint32_t w = 0; w < kKeyWords; ++w) key[w] = sh->dict_keys[g][w];
        slot = group_slot(P.table, key, P.nkeywords, sh->dict_null[g]);
      }
      for (uint32_t e = 0; e <= P.nexprs; ++e) {
        // e == nexprs reduces the row count of the group
        AccT a = Ops::zero();
        uint64_t rows = 0;
        if constexpr (kFastGrouped) {
          constexpr uint32_t NE = SHAPE::Exprs::size;
          if (warp >= PW) {
            if (e < NE) a = slot_value<ACC>(myacc + (g * (NE + 1) + e) * kAccThreads);
            else if (e == P.nexprs) rows = myacc[(g * (NE + 1) + NE) * kAccThreads];
          }
        } else {
#pragma unroll
          for (uint32_t gg = 0; gg < G; ++gg) {
            if (gg == g) {
              rows = grows[gg];
#pragma unroll
              for (uint32_t ee = 0; ee < MAXE; ++ee)
                if (ee == e) a = acc[gg][ee];
            }
          }
        }
        if (warp < PW) { a = Ops::zero(); rows = 0; }
        if (e < P.nexprs) {
#pragma unroll
          for (int o = 16; o; o >>= 1) a = Ops::add(a, Ops::shfl_xor(a, o));
          if (lane == 0 && warp >= PW) *reinterpret_cast<AccT*>(&sh->red[warp - PW][0]) = a;
        } else {
#pragma unroll
          for (int o = 16; o; o >>= 1) rows += __shfl_xor_sync(0xffffffffu, rows, o);
          if (lane == 0 && warp >= PW) sh->red[warp - PW][0] = rows;
        }
        __syncthreads();
        if (threadIdx.x == PW * 32u && slot >= 0) {
          if (e < P.nexprs) {
            AccT t = Ops::zero();
            for (int w = 0; w < kConsumerWarps; ++w) t = Ops::add(t, *reinterpret_cast<AccT*>(&sh->red[w][0]));
            if constexpr (ACC == CLS_F64) {
              // Float64: the CTA's sum goes to its record; the last CTA to finish adds the records in CTA
              // order (below), so the result does not depend on which CTA finishes first
              P.cta_rec[(uint64_t(blockIdx.x) * kRegGroups + g) * (2u + P.nexprs) + 2u + e] = uint64_t(__double_as_longlong(t));
            } else {
              Ops::atomic_add(P.table.acc + (uint64_t(slot) * P.nexprs + e) * P.table.acc_words, t);
            }
          } else {
            uint64_t t = 0;
            for (int w = 0; w < kConsumerWarps; ++w) t += sh->red[w][0];
            if (t) {
              // rows on the fast path had every input valid: they count for every expression
              for (uint32_t ee = 0; ee <= P.nexprs; ++ee)
                atomicAdd(reinterpret_cast<unsigned long long*>(P.table.cnt + uint64_t(slot) * (P.nexprs + 1) + ee), (unsigned long long)t);
            }
          }
        }
        __syncthreads();
      }
      if constexpr (ACC == CLS_F64) {
        if (threadIdx.x == PW * 32u) P.cta_rec[(uint64_t(blockIdx.x) * kRegGroups + g) * (2u + P.nexprs)] = uint64_t(slot) + 1ull;  // 0 = no record
      }
    }
    if constexpr (ACC == CLS_F64) {
      // ===== fixed-order cross-CTA reduction (threadfence reduction: the last CTA to arrive does it) =====
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) sh->last_cta = atomicAdd(P.cta_done, 1u) == gridDim.x - 1u;
      __syncthreads();
      if (sh->last_cta) {
        __threadfence();
        const uint32_t rw = 2u + P.nexprs;                 // record: [slot + 1][unused][sum per argument]
        const uint32_t nrec = gridDim.x * kRegGroups;
        // (staged in the idle stage ring when it is large enough, else read from L2)
        const bool fits = uint64_t(nrec) * rw * 8u <= uint64_t(kNumStages) * P.stage_bytes;
        const uint64_t* recs = fits ? reinterpret_cast<const uint64_t*>(stages) : P.cta_rec;
        if (fits)
          for (uint32_t i = threadIdx.x; i < nrec * rw; i += blockDim.x)
            reinterpret_cast<uint64_t*>(stages)[i] = *reinterpret_cast<volatile uint64_t*>(P.cta_rec + i);
        __syncthreads();
        // Reduction in a FIXED shape (so the bits do not depend on which CTA finished first), in parallel:
        //  1. the distinct slots named by the records of the first CTAs (normally all there are: 1 without GROUP
        //     BY, 4 for Q1) form a short list;
        //  2. one warp per (listed slot, argument): lane l adds the records of CTAs l, l + 32, ... in that order,
        //     then the lanes are combined by a shuffle tree;
        //  3. records naming a slot that is not listed (a group no early CTA met) take the generic path below.
        // What the table already holds came from the rows of the slow path (NULL inputs, more than kRegGroups
        // groups per CTA); every CTA finished those before it took its ticket.
        constexpr uint32_t kListMax = 32;
        uint64_t* lst = reinterpret_cast<uint64_t*>(&sh->red[0][0]);   // kListMax entries: red[] is idle now
        static_assert(sizeof(sh->red) >= kListMax * sizeof(uint64_t), "slot list lives in the reduction scratch");
        if (threadIdx.x == 0) {
          uint32_t n = 0;
          const uint32_t look = min(nrec, 8u * kRegGroups);
          for (uint32_t r = 0; r < look && n < kListMax; ++r) {
            const uint64_t tag = recs[uint64_t(r) * rw];
            bool known = tag == 0ull;
            for (uint32_t q = 0; q < n; ++q) known |= lst[q] == tag;
            if (!known) lst[n++] = tag;
          }
          sh->dict_n = n;   // (the dictionary is no longer needed)
          sh->dict_lock = 0;
        }
        __syncthreads();
        const uint32_t nl = sh->dict_n;
        const uint32_t nwarps = blockDim.x >> 5;
        for (uint32_t task = warp; task < nl * P.nexprs; task += nwarps) {
          const uint32_t e = task % P.nexprs;
          const uint64_t tag = lst[task / P.nexprs];
          double part = 0.0;
          int any = 0;
          for (uint32_t c = lane; c < gridDim.x; c += 32u) {
#pragma unroll
            for (uint32_t g = 0; g < kRegGroups; ++g) {
              const uint64_t* rec = recs + uint64_t(c * kRegGroups + g) * rw;
              if (rec[0] == tag) {
                const double v = __longlong_as_double((long long)rec[2u + e]);
                part = any ? __dadd_rn(part, v) : v;
                any = 1;
              }
            }
          }
#pragma unroll
          for (int o = 16; o; o >>= 1) {
            const double pv = __shfl_down_sync(0xffffffffu, part, o);
            const int pa = __shfl_down_sync(0xffffffffu, any, o);
            if (int(lane) < o && pa) {
              part = any ? __dadd_rn(part, pv) : pv;
              any = 1;
            }
          }
          if (lane == 0 && any) {
            double* dst = reinterpret_cast<double*>(P.table.acc + (tag - 1ull) * P.nexprs + e);
            *dst = __dadd_rn(*dst, part);
          }
        }
        // 3. unlisted slots (rare): a record leads its group if no earlier record names the same slot; (leader,
        // argument) pairs are dealt to the threads, each adds the records of its slot in CTA order
        bool unlisted = false;
        for (uint32_t r = threadIdx.x; r < nrec; r += blockDim.x) {
          const uint64_t tag = recs[uint64_t(r) * rw];
          bool known = tag == 0ull;
          for (uint32_t q = 0; q < nl; ++q) known |= lst[q] == tag;
          unlisted |= !known;
        }
        if (unlisted) atomicExch(&sh->dict_lock, 1u);
        __syncthreads();
        if (sh->dict_lock) {
          for (uint32_t r = threadIdx.x / 16u; r < nrec; r += blockDim.x / 16u) {
            const uint32_t e = threadIdx.x % 16u;
            const uint64_t tag = recs[uint64_t(r) * rw];
            if (tag == 0ull || e >= P.nexprs) continue;
            bool leader = true;
            for (uint32_t q = 0; q < nl; ++q) leader &= lst[q] != tag;
            for (uint32_t q = 0; q < r; ++q) leader &= recs[uint64_t(q) * rw] != tag;
            if (!leader) continue;
            double sum = 0.0;
            bool first = true;
            for (uint32_t q = r; q < nrec; ++q) {
              if (recs[uint64_t(q) * rw] != tag) continue;
              const double v = __longlong_as_double((long long)recs[uint64_t(q) * rw + 2u + e]);
              sum = first ? v : __dadd_rn(sum, v);
              first = false;
            }
            double* dst = reinterpret_cast<double*>(P.table.acc + (tag - 1ull) * P.nexprs + e);
            *dst = __dadd_rn(*dst, sum);
          }
        }
      }
    }
  }
  // counters (RuntimeFilter*/Worker* style metrics)
  {
    uint32_t vals[6] = {n_in, n_bloom, n_filt, n_out, n_bloom_ins, n_bad};
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(P.counters);
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      uint32_t v = vals[q];
#pragma unroll
      for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && v) atomicAdd(dst + q, (unsigned long long)v);
    }
  }
}

}  // namespace pgf
