// The fused scan pipeline kernel (sm_100a).
//
// One persistent CTA per SM streams row tiles of the scan's 64 KiB pages from HBM into a
// 4-stage shared-memory ring with TMA bulk copies (cp.async.bulk + mbarrier transaction
// counts; SASS: UBLKCP), issued by one producer warp.  16 consumer warps evaluate, per
// row and entirely in registers:
//   runtime Bloom probe(s)  -> K2  (pg/backend_service/src/source.rs:496-532)
//   FilterExec predicate     -> K3  (conjunction of <column> <cmp> <literal>)
//   HashJoinExec probe(s)    -> K6  (CollectLeft, Inner; duplicates multiply)
//   ProjectionExec exprs     -> K4  (products of x, (c - x), (c + x))
//   sink: AggregateExec      -> K5  (no-group / register pre-aggregated / global hash)
//         HashJoinExec build [+ RuntimeFilterBuildExec]  -> K6 / K1
// so every page byte is read from HBM exactly once and nothing is materialised.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/pgf_b200.h"
#include "bloom_device.cuh"
#include "device_types.cuh"

namespace pgf {

// Consumer warps per CTA: 16 for streaming sinks; 8 when kRegGroups x kMaxExprs register
// accumulators per thread are live (GROUP BY), which needs the larger register budget.
constexpr int kMaxConsumerWarps = 16;
__host__ __device__ constexpr int consumer_warps(uint32_t sink, bool grouped) { return (sink == SINK_AGG && grouped) ? 8 : 16; }
__host__ __device__ constexpr int pipeline_threads(uint32_t sink, bool grouped) { return (consumer_warps(sink, grouped) + 1) * 32; }
constexpr int kStages = 4;
constexpr uint32_t kAccF64MaxExprs = 8, kAccI128MaxExprs = 6;

struct StageMeta {
  uint32_t nrows;
  uint32_t null_mask;
  uint64_t row_base;
};

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

// Per-row view of the staged tile and of the matched join slots.
struct RowCtx {
  const uint8_t* stage;
  uint32_t r;
  uint32_t tile_nulls;        // page columns with nulls in this tile (already masked by use)
  const uint32_t* pay[kMaxJoins];
  uint32_t occ[kMaxJoins];    // slot occupancy word: bit 0 occupied, bit 1+p payload p is NULL
};

__device__ __forceinline__ bool ref_valid(const DevPlan& P, const RowCtx& c, const DevRef ref) {
  if (ref.src == SRC_PAGE) {
    const DevStageCol& sc = P.scol[ref.idx];
    if (!((c.tile_nulls >> sc.page_col) & 1)) return true;
    return (c.stage[sc.valid_off + (c.r >> 3)] >> (c.r & 7)) & 1;
  }
  const uint32_t occ = ref.src == 1 ? c.occ[0] : c.occ[1];  // static indices: RowCtx stays in registers
  return !((occ >> (1 + (ref.idx >> 8))) & 1);
}

// Raw 128-bit load of a value (low words valid according to the type width).
__device__ __forceinline__ uint4 ref_load(const DevPlan& P, const RowCtx& c, const DevRef ref) {
  uint4 v = make_uint4(0, 0, 0, 0);
  if (ref.src == SRC_PAGE) {
    const DevStageCol& sc = P.scol[ref.idx];
    const uint8_t* p = c.stage + sc.smem_off + c.r * uint32_t(sc.width);
    switch (sc.width) {
      case 16: v = *reinterpret_cast<const uint4*>(p); break;
      case 8: { const uint2 t = *reinterpret_cast<const uint2*>(p); v.x = t.x; v.y = t.y; break; }
      case 4: v.x = *reinterpret_cast<const uint32_t*>(p); break;
      default: v.x = *reinterpret_cast<const uint16_t*>(p); break;
    }
  } else {
    const uint32_t* p = (ref.src == 1 ? c.pay[0] : c.pay[1]) + 3 + (ref.idx & 0xFF);
    const int t = ref.type;
    v.x = __ldg(p);
    if (t == PGF_T_INT64 || t == PGF_T_FLOAT64) v.y = __ldg(p + 1);
    if (t == PGF_T_UTF8VIEW || t == PGF_T_BINARYVIEW || t == PGF_T_DECIMAL128 || t == PGF_T_UUID) {
      v.y = __ldg(p + 1); v.z = __ldg(p + 2); v.w = __ldg(p + 3);
    }
  }
  return v;
}

__device__ __forceinline__ int64_t raw_to_i64(const uint4 v, int type) {
  switch (type) {
    case PGF_T_INT16: return int64_t(int16_t(v.x));
    case PGF_T_INT32: return int64_t(int32_t(v.x));
    default: return int64_t((uint64_t(v.y) << 32) | v.x);
  }
}

// Order-preserving key (signed k0, unsigned k1) of a value: arrow's comparison kernels
// order floats by IEEE totalOrder and strings bytewise; inline views (<= 12 bytes, zero
// padded) compare as (big-endian first 8 bytes, big-endian next 4 bytes, length).
__device__ __forceinline__ bool raw_to_key(const uint4 v, int type, int64_t* k0, uint64_t* k1) {
  switch (type) {
    case PGF_T_INT16: case PGF_T_INT32: case PGF_T_INT64:
      *k0 = raw_to_i64(v, type); *k1 = 0; return true;
    case PGF_T_FLOAT64: {
      int64_t b = int64_t((uint64_t(v.y) << 32) | v.x);
      b ^= int64_t(uint64_t(b >> 63) >> 1);
      *k0 = b; *k1 = 0; return true;
    }
    case PGF_T_FLOAT32: {
      int32_t b = int32_t(v.x);
      b ^= int32_t(uint32_t(b >> 31) >> 1);
      *k0 = b; *k1 = 0; return true;
    }
    case PGF_T_DECIMAL128:
      *k0 = int64_t((uint64_t(v.w) << 32) | v.z); *k1 = (uint64_t(v.y) << 32) | v.x; return true;
    case PGF_T_UTF8VIEW: case PGF_T_BINARYVIEW: {
      const uint64_t hi = (uint64_t(bswap32(v.y)) << 32) | bswap32(v.z);
      *k0 = int64_t(hi ^ 0x8000000000000000ull);
      *k1 = (uint64_t(bswap32(v.w)) << 32) | v.x;
      return v.x <= 12u;  // out-of-line views are not handled by this kernel
    }
    default: return false;
  }
}

__device__ __forceinline__ bool key_cmp(uint32_t cmp, int64_t a0, uint64_t a1, int64_t b0, uint64_t b1) {
  const bool lt = a0 < b0 || (a0 == b0 && a1 < b1);
  const bool eq = a0 == b0 && a1 == b1;
  switch (cmp) {
    case PGF_CMP_LT: return lt;
    case PGF_CMP_LE: return lt || eq;
    case PGF_CMP_GT: return !(lt || eq);
    case PGF_CMP_GE: return !lt;
    case PGF_CMP_EQ: return eq;
    default: return !eq;
  }
}

// ---- expression evaluation ------------------------------------------------------------
template <uint32_t ACC>
__device__ __forceinline__ bool eval_expr(const DevPlan& P, const RowCtx& c, const DevExpr& e,
                                          typename AccOps<ACC>::T* out);

template <>
__device__ __forceinline__ bool eval_expr<CLS_F64>(const DevPlan& P, const RowCtx& c, const DevExpr& e, double* out) {
  double v = 1.0;
  bool ok = true;
#pragma unroll
  for (uint32_t i = 0; i < 3; ++i) {
    if (i < e.nfactors) {
      const DevFactor& f = e.f[i];
      ok &= ref_valid(P, c, f.ref);
      const uint4 raw = ref_load(P, c, f.ref);
      double x;
      if (f.ref.type == PGF_T_FLOAT64) x = __longlong_as_double((long long)((uint64_t(raw.y) << 32) | raw.x));
      else if (f.ref.type == PGF_T_FLOAT32) x = double(__uint_as_float(raw.x));
      else x = double(raw_to_i64(raw, f.ref.type));  // AVG over integers runs on the Float64 cast
      // one IEEE operation per node, no FMA contraction (explicit _rn intrinsics)
      if (f.kind == PGF_FACTOR_CONST_MINUS_COL) x = __dsub_rn(f.cf, x);
      else if (f.kind == PGF_FACTOR_CONST_PLUS_COL) x = __dadd_rn(f.cf, x);
      v = i == 0 ? x : __dmul_rn(v, x);
    }
  }
  *out = v;
  return ok;
}

template <>
__device__ __forceinline__ bool eval_expr<CLS_I64>(const DevPlan& P, const RowCtx& c, const DevExpr& e, uint64_t* out) {
  uint64_t v = 1;
  bool ok = true;
#pragma unroll
  for (uint32_t i = 0; i < 3; ++i) {
    if (i < e.nfactors) {
      const DevFactor& f = e.f[i];
      ok &= ref_valid(P, c, f.ref);
      uint64_t x = uint64_t(raw_to_i64(ref_load(P, c, f.ref), f.ref.type));
      if (f.kind == PGF_FACTOR_CONST_MINUS_COL) x = uint64_t(f.ci_lo) - x;
      else if (f.kind == PGF_FACTOR_CONST_PLUS_COL) x = uint64_t(f.ci_lo) + x;
      v = i == 0 ? x : v * x;  // wrapping
    }
  }
  *out = v;
  return ok;
}

template <>
__device__ __forceinline__ bool eval_expr<CLS_I128>(const DevPlan& P, const RowCtx& c, const DevExpr& e, I128* out) {
  I128 v{1, 0};
  bool ok = true;
#pragma unroll
  for (uint32_t i = 0; i < 3; ++i) {
    if (i < e.nfactors) {
      const DevFactor& f = e.f[i];
      ok &= ref_valid(P, c, f.ref);
      const uint4 raw = ref_load(P, c, f.ref);
      I128 x;
      if (f.ref.type == PGF_T_DECIMAL128) {
        x.lo = (uint64_t(raw.y) << 32) | raw.x;
        x.hi = (uint64_t(raw.w) << 32) | raw.z;
      } else {
        const int64_t s = raw_to_i64(raw, f.ref.type);
        x.lo = uint64_t(s);
        x.hi = s < 0 ? ~0ull : 0ull;
      }
      const I128 cst{uint64_t(f.ci_lo), uint64_t(f.ci_hi)};
      if (f.kind == PGF_FACTOR_CONST_MINUS_COL) x = i128_sub(cst, x);
      else if (f.kind == PGF_FACTOR_CONST_PLUS_COL) x = i128_add(cst, x);
      v = i == 0 ? x : i128_mul(v, x);
    }
  }
  *out = v;
  return ok;
}

// ---- global group table ---------------------------------------------------------------
// Returns the slot holding `key` (inserting it if absent) or -1 when the table is full.
__device__ __forceinline__ int64_t group_slot(const GroupTable& t, const uint64_t* key, uint32_t nwords, uint32_t knull) {
  uint64_t h = 0x9E3779B97F4A7C15ull ^ knull;
#pragma unroll
  for (uint32_t w = 0; w < kKeyWords; ++w)
    if (w < nwords) h = mix64(h ^ key[w]) + 0x632BE59BD9B4E019ull;
  uint32_t i = uint32_t(h) & t.mask;
  const uint32_t ready = 2u | (knull << 8);
  for (uint32_t probes = 0; probes <= t.mask; ++probes, i = (i + 1) & t.mask) {
    uint32_t s = *reinterpret_cast<volatile uint32_t*>(t.state + i);
    if (s == 0) {
      const uint32_t old = atomicCAS(t.state + i, 0u, 1u);
      if (old == 0) {
#pragma unroll
        for (uint32_t w = 0; w < kKeyWords; ++w)
          if (w < nwords) t.keys[uint64_t(i) * kKeyWords + w] = key[w];
        __threadfence();
        atomicExch(t.state + i, ready);
        atomicAdd(t.used, 1u);
        return i;
      }
      s = old;
    }
    while ((s & 3u) == 1u) s = *reinterpret_cast<volatile uint32_t*>(t.state + i);  // being published
    if (s == ready) {
      __threadfence();
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
  uint32_t dict_null[kRegGroups];
  uint32_t dict_n;
  uint32_t dict_lock;
  // block reduction scratch
  uint64_t red[kMaxConsumerWarps][2];
};

__device__ __forceinline__ int dict_find(const BlockShared* sh, const uint64_t* key, uint32_t nwords, uint32_t knull, uint32_t n) {
  for (uint32_t g = 0; g < n; ++g) {
    bool same = *reinterpret_cast<const volatile uint32_t*>(&sh->dict_null[g]) == knull;
#pragma unroll
    for (uint32_t w = 0; w < kKeyWords; ++w)
      if (w < nwords) same &= *reinterpret_cast<const volatile uint64_t*>(&sh->dict_keys[g][w]) == key[w];
    if (same) return int(g);
  }
  return -1;
}

// Find the key in the CTA dictionary, appending it while there is room.  Returns -1 when
// the key is not one of the (at most kRegGroups) register-resident groups.
__device__ __forceinline__ int dict_lookup_or_insert(BlockShared* sh, const uint64_t* key, uint32_t nwords, uint32_t knull) {
  uint32_t n = *reinterpret_cast<volatile uint32_t*>(&sh->dict_n);
  int g = dict_find(sh, key, nwords, knull, n);
  if (g >= 0 || n == kRegGroups) return g;
  // rare path: append under a CTA-wide spin lock (at most kRegGroups successful appends)
  bool done = false;
  while (!done) {
    if (atomicCAS(&sh->dict_lock, 0u, 1u) == 0u) {
      n = *reinterpret_cast<volatile uint32_t*>(&sh->dict_n);
      g = dict_find(sh, key, nwords, knull, n);
      if (g < 0 && n < kRegGroups) {
#pragma unroll
        for (uint32_t w = 0; w < kKeyWords; ++w)
          if (w < nwords) sh->dict_keys[n][w] = key[w];
        sh->dict_null[n] = knull;
        __threadfence_block();
        *reinterpret_cast<volatile uint32_t*>(&sh->dict_n) = n + 1;
        g = int(n);
      }
      __threadfence_block();
      atomicExch(&sh->dict_lock, 0u);
      done = true;
    }
  }
  return g;
}

// ---- the kernel ----------------------------------------------------------------------
template <uint32_t SINK, uint32_t ACC, bool GROUPED, uint32_t NJ>
__global__ void __launch_bounds__(pipeline_threads(SINK, GROUPED), 1) pipeline_kernel(const __grid_constant__ DevPlan P) {
  constexpr int kConsumerWarps = consumer_warps(SINK, GROUPED);
  using Ops = AccOps<ACC>;
  using AccT = typename Ops::T;
  constexpr uint32_t MAXE = SINK == SINK_AGG ? Ops::kMaxExprs : 1;
  constexpr uint32_t G = (SINK == SINK_AGG && GROUPED) ? kRegGroups : 1;

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
  uint32_t n_in = 0, n_bloom = 0, n_filt = 0, n_out = 0, n_bad = 0, n_bloom_ins = 0;

  if (warp == 0) {
    // ===== producer: TMA bulk copies of the needed column slices of each row tile =====
    uint32_t k = 0;
    for (uint32_t item = blockIdx.x; item < P.nitems; item += gridDim.x, ++k) {
      const uint32_t s = k % kStages;
      const uint32_t page = item / P.tiles_per_page, tile = item - page * P.tiles_per_page;
      const PageDesc d = P.descs[page];
      const LayoutClass* lc = P.classes + d.layout_class;
      const uint32_t r0 = tile * P.tile_rows;
      const uint32_t n = d.row_count > r0 ? min(d.row_count - r0, P.tile_rows) : 0u;
      const uint8_t* page_base = P.pages + uint64_t(page) * P.page_stride;
      uint8_t* stage = stages + size_t(s) * P.stage_bytes;
      uint32_t bytes = 0;
      const uint8_t* src = nullptr;
      uint8_t* dst = nullptr;
      if (n) {
        if (lane < P.nstage_cols) {
          const DevStageCol sc = P.scol[lane];
          bytes = (n * sc.width + 15u) & ~15u;
          src = page_base + lc->values_off[sc.page_col] + r0 * uint32_t(sc.width);
          dst = stage + sc.smem_off;
        } else if (lane >= 16 && lane - 16 < P.nstage_cols) {
          const DevStageCol sc = P.scol[lane - 16];
          if (sc.nullable && ((d.null_mask & P.used_null_mask) >> sc.page_col) & 1) {
            bytes = (((n + 7u) >> 3) + 15u) & ~15u;
            src = page_base + lc->validity_off[sc.page_col] + (r0 >> 3);
            dst = stage + sc.valid_off;
          }
        }
      }
      uint32_t total = bytes;
#pragma unroll
      for (int o = 16; o; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
      mbar_wait(&sh->empty[s], ((k / kStages) & 1u) ^ 1u);
      if (lane == 0) {
        sh->meta[s].nrows = n;
        sh->meta[s].null_mask = d.null_mask & P.used_null_mask;
        sh->meta[s].row_base = d.row_base + r0;
        mbar_arrive_expect_tx(&sh->full[s], total);
      }
      __syncwarp();
      if (bytes) tma_load_1d(dst, src, bytes, &sh->full[s]);
    }
  } else {
    // ===== consumers =====
    const uint32_t ct = threadIdx.x - 32;
    constexpr uint32_t NCT = kConsumerWarps * 32;
    uint32_t k = 0;
    for (uint32_t item = blockIdx.x; item < P.nitems; item += gridDim.x, ++k) {
      const uint32_t s = k % kStages;
      mbar_wait(&sh->full[s], (k / kStages) & 1u);
      const uint32_t nrows = sh->meta[s].nrows;
      RowCtx c;
      c.stage = stages + size_t(s) * P.stage_bytes;
      c.tile_nulls = sh->meta[s].null_mask;
      for (uint32_t r = ct; r < nrows; r += NCT) {
        c.r = r;
        ++n_in;
        bool keep = true;
        // -- runtime Bloom probes: NULL key => DefinitelyAbsent (shared.rs:367-374)
        for (uint32_t b = 0; b < P.nbloom && keep; ++b) {
          const DevBloomProbe& bp = P.bloom[b];
          keep = ref_valid(P, c, bp.key) &&
                 bloom_contains(bp.bloom, uint64_t(raw_to_i64(ref_load(P, c, bp.key), bp.key.type)));
        }
        if (!keep) continue;
        ++n_bloom;
        // -- FilterExec: every conjunct must be TRUE (NULL drops the row)
#pragma unroll
        for (uint32_t t = 0; t < kMaxTerms; ++t) {
          if (t < P.nterms) {
            const DevTerm& tm = P.terms[t];
            int64_t k0;
            uint64_t k1;
            const bool ok = raw_to_key(ref_load(P, c, tm.ref), tm.ref.type, &k0, &k1);
            const bool valid = ref_valid(P, c, tm.ref);
            if (valid && !ok) ++n_bad;
            keep &= valid && ok && key_cmp(tm.cmp, k0, k1, tm.k0, tm.k1);
          }
        }
        if (!keep) continue;
        ++n_filt;

        // -- sink (optionally behind one HashJoinExec probe)
        auto sink = [&](const RowCtx& rc) {
          ++n_out;
          if constexpr (SINK == SINK_AGG) {
            AccT v[MAXE];
            bool all_valid = true;
            uint32_t valid_mask = 0;
#pragma unroll
            for (uint32_t e = 0; e < MAXE; ++e) {
              v[e] = Ops::zero();
              if (e < P.nexprs) {
                const bool ok = eval_expr<ACC>(P, rc, P.exprs[e], &v[e]);
                all_valid &= ok;
                valid_mask |= uint32_t(ok) << e;
              }
            }
            uint64_t key[kKeyWords] = {0, 0, 0, 0};
            uint32_t knull = 0;
            int g = 0;
            if constexpr (GROUPED) {
#pragma unroll
              for (uint32_t kp = 0; kp < 4; ++kp) {
                if (kp < P.nkeys) {
                  const DevKeyPart& part = P.keys[kp];
                  if (!ref_valid(P, rc, part.ref)) {
                    knull |= 1u << kp;  // NULL keys form one group; words stay zero
                  } else {
                    const uint4 raw = ref_load(P, rc, part.ref);
                    uint64_t w0, w1 = 0;
                    if (part.nwords == 1) {
                      w0 = uint64_t(raw_to_i64(raw, part.ref.type));
                    } else {
                      if ((part.ref.type == PGF_T_UTF8VIEW || part.ref.type == PGF_T_BINARYVIEW) && raw.x > 12u) ++n_bad;
                      w0 = (uint64_t(raw.y) << 32) | raw.x;
                      w1 = (uint64_t(raw.w) << 32) | raw.z;
                    }
#pragma unroll
                    for (uint32_t w = 0; w < kKeyWords; ++w) {  // static indices keep key[] in registers
                      if (w == part.word) key[w] = w0;
                      if (part.nwords == 2 && w == part.word + 1u) key[w] = w1;
                    }
                  }
                }
              }
              g = dict_lookup_or_insert(sh, key, P.nkeywords, knull);
            }
            if (all_valid && g >= 0) {
              // fast path: register accumulators (predicated over the register groups)
#pragma unroll
              for (uint32_t gg = 0; gg < G; ++gg) {
                if (G == 1 || gg == uint32_t(g)) {
                  grows[gg] += 1;
#pragma unroll
                  for (uint32_t e = 0; e < MAXE; ++e)
                    if (e < P.nexprs) acc[gg][e] = Ops::add(acc[gg][e], v[e]);
                }
              }
            } else {
              // slow path: straight to the global table (NULL inputs, or > kRegGroups groups)
              const int64_t slot = GROUPED ? group_slot(P.table, key, P.nkeywords, knull) : 0;
              if (slot >= 0) {
#pragma unroll
                for (uint32_t e = 0; e < MAXE; ++e) {
                  if (e < P.nexprs && ((valid_mask >> e) & 1)) {
                    Ops::atomic_add(P.table.acc + (uint64_t(slot) * P.nexprs + e) * P.table.acc_words, v[e]);
                    atomicAdd(reinterpret_cast<unsigned long long*>(P.table.cnt + uint64_t(slot) * (P.nexprs + 1) + e), 1ull);
                  }
                }
                atomicAdd(reinterpret_cast<unsigned long long*>(P.table.cnt + uint64_t(slot) * (P.nexprs + 1) + P.nexprs), 1ull);
              }
            }
          } else if constexpr (SINK == SINK_JOIN_BUILD) {
            const JoinBuild& jb = P.build;
            if (ref_valid(P, rc, jb.key)) {  // NULL keys never match: not inserted
              const int64_t key = raw_to_i64(ref_load(P, rc, jb.key), jb.key.type);
              uint32_t pay[5] = {0, 0, 0, 0, 0};
              uint32_t occ = 1u;
#pragma unroll
              for (uint32_t p = 0; p < 4; ++p) {
                if (p >= jb.npayload) continue;
                if (!ref_valid(P, rc, jb.payload[p])) { occ |= 2u << p; continue; }
                const uint4 raw = ref_load(P, rc, jb.payload[p]);
                const uint32_t w = jb.payload_word[p], nw = jb.payload_nwords[p];
#pragma unroll
                for (uint32_t q = 0; q < 5; ++q) {  // static indices keep pay[] in registers
                  if (q == w) pay[q] = raw.x;
                  if (nw >= 2 && q == w + 1) pay[q] = raw.y;
                  if (nw == 4 && q == w + 2) pay[q] = raw.z;
                  if (nw == 4 && q == w + 3) pay[q] = raw.w;
                }
              }
              uint32_t i = uint32_t(mix64(uint64_t(key))) & jb.mask;
              for (;;) {  // capacity >= 2 x rows: an empty slot always exists
                uint32_t* slot = reinterpret_cast<uint32_t*>(jb.slots + uint64_t(i) * jb.slot_u4);
                if (atomicCAS(slot + 2, 0u, occ) == 0u) {
                  slot[0] = uint32_t(uint64_t(key));
                  slot[1] = uint32_t(uint64_t(key) >> 32);
                  slot[3] = pay[0];
                  if (jb.slot_u4 == 2) { slot[4] = pay[1]; slot[5] = pay[2]; slot[6] = pay[3]; slot[7] = pay[4]; }
                  break;
                }
                i = (i + 1) & jb.mask;
              }
              if (P.has_build_bloom) { bloom_insert(P.build_bloom, uint64_t(key)); ++n_bloom_ins; }
            }
          }
        };

        if constexpr (NJ == 0) {
          c.pay[0] = nullptr; c.occ[0] = 0;
          sink(c);
        } else {
          const DevJoin& j = P.joins[0];
          if (!ref_valid(P, c, j.key)) continue;  // NULL keys never match
          const int64_t key = raw_to_i64(ref_load(P, c, j.key), j.key.type);
          const uint32_t klo = uint32_t(uint64_t(key)), khi = uint32_t(uint64_t(key) >> 32);
          uint32_t i = uint32_t(mix64(uint64_t(key))) & j.mask;
          for (;;) {
            const uint4* slot = j.slots + uint64_t(i) * j.slot_u4;
            const uint4 s0 = __ldg(slot);
            if ((s0.z & 1u) == 0u) break;
            if (s0.x == klo && s0.y == khi) {
              c.pay[0] = reinterpret_cast<const uint32_t*>(slot);
              c.occ[0] = s0.z;
              sink(c);
            }
            i = (i + 1) & j.mask;
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->empty[s]);
    }
  }
  __syncthreads();

  // ===== epilogue: block-reduce the register accumulators, then one atomic per CTA =====
  if constexpr (SINK == SINK_AGG) {
    const uint32_t ng = GROUPED ? sh->dict_n : 1u;
    for (uint32_t g = 0; g < ng; ++g) {
      int64_t slot = 0;
      if (GROUPED && threadIdx.x == 32) {
        uint64_t key[kKeyWords];
#pragma unroll
        for (uint32_t w = 0; w < kKeyWords; ++w) key[w] = sh->dict_keys[g][w];
        slot = group_slot(P.table, key, P.nkeywords, sh->dict_null[g]);
      }
      for (uint32_t e = 0; e <= P.nexprs; ++e) {
        // e == nexprs reduces the row count of the group
        AccT a = Ops::zero();
        uint64_t rows = 0;
#pragma unroll
        for (uint32_t gg = 0; gg < G; ++gg) {
          if (gg == g) {
            rows = grows[gg];
#pragma unroll
            for (uint32_t ee = 0; ee < MAXE; ++ee)
              if (ee == e) a = acc[gg][ee];
          }
        }
        if (warp == 0) { a = Ops::zero(); rows = 0; }
        if (e < P.nexprs) {
#pragma unroll
          for (int o = 16; o; o >>= 1) a = Ops::add(a, Ops::shfl_xor(a, o));
          if (lane == 0 && warp > 0) *reinterpret_cast<AccT*>(&sh->red[warp - 1][0]) = a;
        } else {
#pragma unroll
          for (int o = 16; o; o >>= 1) rows += __shfl_xor_sync(0xffffffffu, rows, o);
          if (lane == 0 && warp > 0) sh->red[warp - 1][0] = rows;
        }
        __syncthreads();
        if (threadIdx.x == 32 && slot >= 0) {
          if (e < P.nexprs) {
            AccT t = Ops::zero();
            for (int w = 0; w < kConsumerWarps; ++w) t = Ops::add(t, *reinterpret_cast<AccT*>(&sh->red[w][0]));
            Ops::atomic_add(P.table.acc + (uint64_t(slot) * P.nexprs + e) * P.table.acc_words, t);
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
