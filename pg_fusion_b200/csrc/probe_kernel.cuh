// The fused scan pipeline for selective plans: anything with a HashJoinExec probe or a
// HashJoinExec build sink (sm_100a).
//
// Same front end as pipeline_kernel.cuh (one persistent CTA per SM, one producer warp feeding a
// shared-memory ring with TMA bulk copies + mbarrier transaction counts), but the consumer side is
// organised around *selection-vector compaction* instead of per-thread predication:
//
//   stage A  filter      28 warps walk the staged tile 32 rows at a time: runtime Bloom probes and the
//                        FilterExec conjuncts in registers, then ballot + prefix-popcount compaction of
//                        the surviving (page, row, join key) triples into a per-warp queue in shared
//                        memory.  Only the predicate columns and the probe key are staged at all.
//   stage B  tag probe   whenever a warp's queue holds 32 survivors it pops them onto DENSE lanes,
//                        hashes the keys and issues one aligned 8-byte load of the home bucket's tag
//                        bytes (L2 resident directory, evict_last).  The loads stay in flight while
//                        the warp goes back to stage A; they are resolved (SIMD-in-register byte
//                        compares) just before the next batch is issued.  Tag hits are compacted into
//                        a second per-warp queue.
//   stage C  match+sink  32 tag hits at a time, again on dense lanes: walk the probe chain, compare
//                        the slot keys (duplicates multiply, NULL keys never match), optionally probe
//                        a second join table, and feed the sink.  Columns that only matched rows need
//                        (aggregate arguments, group keys, build payloads) are read straight from the
//                        page in HBM here -- late materialisation: for TPC-H Q3's lineitem side that is
//                        1 % of the rows, so 16 of the 36 algorithmic bytes per row never leave DRAM.
//
// Sinks: AggregateExec into the global group table (grouped) or warp-reduced into slot 0 (no GROUP
// BY); HashJoinExec build side as a dense array of build rows (warp-aggregated append; the table is
// built from the rows afterwards at exactly the capacity the row count asks for, see
// join_build_kernel in pipeline.cu) [+ RuntimeFilterBuildExec]; row count.
#pragma once
#include "pipeline_kernel.cuh"

namespace pgf {

#ifndef PGF_PROBE_WARPS
#define PGF_PROBE_WARPS 20
#endif
constexpr int kPConsumerWarps = PGF_PROBE_WARPS;
constexpr int kPThreads = (kPConsumerWarps + 1) * 32;
constexpr uint32_t kPMaxStages = 6;
constexpr uint32_t kPQueueEntries = 64;                              // per warp and queue; drained 32 at a time
constexpr uint32_t kPQueueBytesPerWarp = 2u * kPQueueEntries * 16u;  // survivors + tag hits

struct PStageMeta {
  uint32_t nrows, null_mask, page, r0;
};
struct ProbeShared {
  uint64_t full[kPMaxStages];
  uint64_t empty[kPMaxStages];
  PStageMeta meta[kPMaxStages];
};
__host__ __device__ constexpr uint32_t probe_shared_bytes() { return uint32_t((sizeof(ProbeShared) + 127) & ~size_t(127)); }

// ---- L2 residency control: the page stream is read once (evict_first), the tag directories and the
// group table are what should stay (evict_last)
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_load_1d_hint(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
}
// the 8 tag bytes of one bucket (8-byte aligned), predicated so that idle lanes issue nothing
__device__ __forceinline__ uint2 ldg_tags8(const uint8_t* p, bool pred, uint64_t policy) {
  uint2 v = make_uint2(0u, 0u);
  asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p ld.global.nc.L2::cache_hint.v2.u32 {%0, %1}, [%2], %4;\n\t}"
      : "+r"(v.x), "+r"(v.y)
      : "l"(p), "r"(uint32_t(pred)), "l"(policy));
  return v;
}

// 0x80 in every byte of x that is zero (exact: no carries between bytes)
__device__ __forceinline__ uint32_t zero_bytes(uint32_t x) { return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu); }

// One bucket window: `cand` has bit 8k+7 set for every tag byte k that equals `tag` and lies before the
// first empty tag; `ended` says the window holds an empty tag (the probe chain stops here).
struct TagWindow {
  uint64_t cand;
  bool ended;
};
__device__ __forceinline__ TagWindow scan_tags(uint2 w, uint32_t tag) {
  const uint32_t t4 = tag * 0x01010101u;
  const uint64_t mz = (uint64_t(zero_bytes(w.y)) << 32) | zero_bytes(w.x);
  const uint64_t mt = (uint64_t(zero_bytes(w.y ^ t4)) << 32) | zero_bytes(w.x ^ t4);
  TagWindow r;
  r.ended = mz != 0ull;
  r.cand = mt & (r.ended ? ((mz & (0ull - mz)) - 1ull) : ~0ull);
  return r;
}

// ---- a row of a page in HBM (stage C reads what it needs from there) ----------------------
struct GRow {
  const uint8_t* page;
  const LayoutClass* lc;
  uint32_t r;
  uint32_t nulls;                  // page columns with nulls in this page (already masked by use)
  const uint32_t *pay0, *pay1;     // matched join slots (u32 words: key lo, key hi, occ, payload...)
  uint32_t occ0, occ1;
  __device__ __forceinline__ const uint32_t* pay(uint32_t src) const { return src == 1 ? pay0 : pay1; }
  __device__ __forceinline__ uint32_t occ(uint32_t src) const { return src == 1 ? occ0 : occ1; }
};

__device__ __forceinline__ bool g_valid(const DevRef& ref, const GRow& g) {
  if (ref.src == SRC_PAGE) {
    if (ref.valid_off == kNoValidity || !((g.nulls >> ref.pcol) & 1)) return true;
    return (g.page[g.lc->validity_off[ref.pcol] + (g.r >> 3)] >> (g.r & 7)) & 1;
  }
  return !((g.occ(ref.src) >> (1 + ref.pcol)) & 1);
}
__device__ __forceinline__ int64_t g_i64(const DevRef& ref, const GRow& g) {
  if (ref.src == SRC_PAGE) {
    const uint8_t* p = g.page + g.lc->values_off[ref.pcol];
    switch (ref.ld) {
      case LD_I16: return int64_t(reinterpret_cast<const int16_t*>(p)[g.r]);
      case LD_I32: return int64_t(reinterpret_cast<const int32_t*>(p)[g.r]);
      default: return reinterpret_cast<const int64_t*>(p)[g.r];
    }
  }
  const uint32_t* p = g.pay(ref.src) + 3 + ref.off;
  switch (ref.ld) {
    case LD_I16: return int64_t(int16_t(__ldg(p)));
    case LD_I32: return int64_t(int32_t(__ldg(p)));
    default: return int64_t((uint64_t(__ldg(p + 1)) << 32) | __ldg(p));
  }
}
__device__ __forceinline__ uint4 g_u128(const DevRef& ref, const GRow& g) {
  if (ref.src == SRC_PAGE) return reinterpret_cast<const uint4*>(g.page + g.lc->values_off[ref.pcol])[g.r];
  const uint32_t* p = g.pay(ref.src) + 3 + ref.off;
  return make_uint4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
}
__device__ __forceinline__ double g_f64(const DevRef& ref, const GRow& g) {
  if (ref.src == SRC_PAGE) {
    const uint8_t* p = g.page + g.lc->values_off[ref.pcol];
    if (ref.ld == LD_F64) return reinterpret_cast<const double*>(p)[g.r];
    if (ref.ld == LD_F32) return double(reinterpret_cast<const float*>(p)[g.r]);
    return double(g_i64(ref, g));  // AVG over integers runs on the Float64 cast
  }
  const uint32_t* p = g.pay(ref.src) + 3 + ref.off;
  if (ref.ld == LD_F64) return __longlong_as_double((long long)((uint64_t(__ldg(p + 1)) << 32) | __ldg(p)));
  if (ref.ld == LD_F32) return double(__uint_as_float(__ldg(p)));
  return double(g_i64(ref, g));
}

// ProjectionExec / aggregate arguments over a GRow: one IEEE operation per node (never an FMA), wrapping
// i64 / i128 -- the same arithmetic as eval_expr_* in pipeline_kernel.cuh.
template <uint32_t ACC>
__device__ __forceinline__ typename AccOps<ACC>::T g_eval(const DevExpr& e, const GRow& g) {
  if constexpr (ACC == CLS_F64) {
    double v = 1.0;
#pragma unroll
    for (uint32_t i = 0; i < 3; ++i) {
      if (i < e.nfactors) {
        const DevFactor& f = e.f[i];
        double x = g_f64(f.ref, g);
        if (f.kind == PGF_FACTOR_CONST_MINUS_COL) x = __dsub_rn(f.cf, x);
        else if (f.kind == PGF_FACTOR_CONST_PLUS_COL) x = __dadd_rn(f.cf, x);
        v = i == 0 ? x : __dmul_rn(v, x);
      }
    }
    return v;
  } else if constexpr (ACC == CLS_I64) {
    uint64_t v = 1;
#pragma unroll
    for (uint32_t i = 0; i < 3; ++i) {
      if (i < e.nfactors) {
        const DevFactor& f = e.f[i];
        uint64_t x = uint64_t(g_i64(f.ref, g));
        if (f.kind == PGF_FACTOR_CONST_MINUS_COL) x = uint64_t(f.ci_lo) - x;
        else if (f.kind == PGF_FACTOR_CONST_PLUS_COL) x = uint64_t(f.ci_lo) + x;
        v = i == 0 ? x : v * x;
      }
    }
    return v;
  } else {
    I128 v{1, 0};
#pragma unroll
    for (uint32_t i = 0; i < 3; ++i) {
      if (i < e.nfactors) {
        const DevFactor& f = e.f[i];
        I128 x;
        if (f.ref.ld == LD_DEC) {
          const uint4 raw = g_u128(f.ref, g);
          x.lo = (uint64_t(raw.y) << 32) | raw.x;
          x.hi = (uint64_t(raw.w) << 32) | raw.z;
        } else {
          const int64_t s = g_i64(f.ref, g);
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
}
__device__ __forceinline__ bool g_expr_valid(const DevExpr& e, const GRow& g) {
  if (!(g.nulls & e.null_cols) && !e.has_payload) return true;
  bool ok = true;
#pragma unroll
  for (uint32_t i = 0; i < 3; ++i)
    if (i < e.nfactors) ok &= g_valid(e.f[i].ref, g);
  return ok;
}

// ---- probe chain iterator (stage C) --------------------------------------------------------
struct JoinIter {
  uint64_t cand;
  uint32_t base, tag, klo, khi;
  bool ended;

  __device__ __forceinline__ void load(const DevJoin& j) {
    const uint2 w = __ldg(reinterpret_cast<const uint2*>(j.tags + base));
    const TagWindow tw = scan_tags(w, tag);
    cand = tw.cand;
    ended = tw.ended;
  }
  __device__ __forceinline__ void init(const DevJoin& j, int64_t key, bool valid) {
    const uint64_t h = join_hash(key);
    base = join_home(h, j.shift);
    tag = join_tag8(h, j.shift);
    klo = uint32_t(uint64_t(key));
    khi = uint32_t(uint64_t(key) >> 32);
    cand = 0;
    ended = true;
    if (valid) load(j);  // NULL keys never match
  }
  // next slot of the chain whose key equals the probe key, or nullptr
  __device__ __forceinline__ const uint32_t* next(const DevJoin& j, uint32_t* occ) {
    for (;;) {
      while (cand) {
        const uint32_t b = uint32_t(__ffsll((long long)cand)) - 1u;
        cand &= cand - 1ull;
        const uint4* slot = j.slots + uint64_t(base + (b >> 3)) * j.slot_u4;
        const uint4 s0 = __ldg(slot);
        if (s0.x == klo && s0.y == khi) {
          *occ = s0.z;
          return reinterpret_cast<const uint32_t*>(slot);
        }
      }
      if (ended) return nullptr;
      base = (base + kJoinBucket) & j.mask;
      load(j);
    }
  }
};

// ---- sinks (stage C, dense lanes) -----------------------------------------------------------
template <uint32_t ACC>
__device__ __forceinline__ void sink_agg_grouped(const DevPlan& P, const GRow& g, uint32_t& bad) {
  using Ops = AccOps<ACC>;
  uint64_t key[kKeyWords] = {0, 0, 0, 0};
  uint32_t knull = 0;
#pragma unroll
  for (uint32_t kp = 0; kp < 4; ++kp) {
    if (kp < P.nkeys) {
      const DevKeyPart& part = P.keys[kp];
      if (!g_valid(part.ref, g)) {
        knull |= 1u << kp;  // NULL keys form one group; words stay zero
      } else {
        uint64_t w0, w1 = 0;
        if (part.nwords == 1) {
          w0 = uint64_t(g_i64(part.ref, g));
        } else {
          const uint4 raw = g_u128(part.ref, g);
          if (part.ref.ld == LD_VIEW && raw.x > 12u) ++bad;
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
  const int64_t slot = group_slot(P.table, key, P.nkeywords, knull);
  if (slot < 0) return;
  for (uint32_t e = 0; e < P.nexprs; ++e) {
    if (g_expr_valid(P.exprs[e], g)) {
      Ops::atomic_add(P.table.acc + (uint64_t(slot) * P.nexprs + e) * P.table.acc_words, g_eval<ACC>(P.exprs[e], g));
      atomicAdd(reinterpret_cast<unsigned long long*>(P.table.cnt + uint64_t(slot) * (P.nexprs + 1) + e), 1ull);
    }
  }
  atomicAdd(reinterpret_cast<unsigned long long*>(P.table.cnt + uint64_t(slot) * (P.nexprs + 1) + P.nexprs), 1ull);
}

// no GROUP BY: the matches of the warp are reduced with shuffles, one atomic per argument and warp
template <uint32_t ACC>
__device__ __forceinline__ void sink_agg_single(const DevPlan& P, const GRow& g, bool found, uint32_t lane) {
  using Ops = AccOps<ACC>;
  for (uint32_t e = 0; e < P.nexprs; ++e) {
    const bool ok = found && g_expr_valid(P.exprs[e], g);
    typename Ops::T v = Ops::zero();
    if (ok) v = g_eval<ACC>(P.exprs[e], g);
    const uint32_t n = __popc(__ballot_sync(0xffffffffu, ok));
#pragma unroll
    for (int o = 16; o; o >>= 1) v = Ops::add(v, Ops::shfl_xor(v, o));
    if (lane == 0 && n) {
      Ops::atomic_add(P.table.acc + uint64_t(e) * P.table.acc_words, v);
      atomicAdd(reinterpret_cast<unsigned long long*>(P.table.cnt + e), (unsigned long long)n);
    }
  }
  const uint32_t rows = __popc(__ballot_sync(0xffffffffu, found));
  if (lane == 0 && rows) atomicAdd(reinterpret_cast<unsigned long long*>(P.table.cnt + P.nexprs), (unsigned long long)rows);
}

// HashJoinExec build side: append {key, occupancy / NULL flags, payload} to the dense row array
__device__ __forceinline__ void sink_build(const DevPlan& P, const GRow& g, bool found, uint32_t lane, uint32_t& n_bloom_ins) {
  const JoinBuild& jb = P.build;
  const bool valid = found && g_valid(jb.key, g);  // NULL keys never match: not inserted
  const uint32_t m = __ballot_sync(0xffffffffu, valid);
  if (!m) return;
  unsigned long long base = 0;
  if (lane == 0) base = atomicAdd(P.build_count, (unsigned long long)__popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (!valid) return;
  const unsigned long long pos = base + __popc(m & ((1u << lane) - 1u));
  const int64_t key = g_i64(jb.key, g);
  uint32_t pay[5] = {0, 0, 0, 0, 0};
  uint32_t occ = 1u;
#pragma unroll
  for (uint32_t p = 0; p < 4; ++p) {
    if (p >= jb.npayload) continue;
    if (!g_valid(jb.payload[p], g)) { occ |= 2u << p; continue; }
    const uint32_t w = jb.payload_word[p], nw = jb.payload_nwords[p];
    uint4 raw = make_uint4(0, 0, 0, 0);
    if (nw == 4) raw = g_u128(jb.payload[p], g);
    else {
      const int64_t x = jb.payload[p].ld == LD_F64 ? __double_as_longlong(g_f64(jb.payload[p], g))
                      : jb.payload[p].ld == LD_F32 ? int64_t(__float_as_uint(float(g_f64(jb.payload[p], g))))
                                                   : g_i64(jb.payload[p], g);
      raw.x = uint32_t(uint64_t(x));
      raw.y = uint32_t(uint64_t(x) >> 32);
    }
#pragma unroll
    for (uint32_t q = 0; q < 5; ++q) {  // static indices keep pay[] in registers
      if (q == w) pay[q] = raw.x;
      if (nw >= 2 && q == w + 1) pay[q] = raw.y;
      if (nw == 4 && q == w + 2) pay[q] = raw.z;
      if (nw == 4 && q == w + 3) pay[q] = raw.w;
    }
  }
  if (pos < jb.rows_cap) {
    uint4* row = jb.rows + pos * jb.slot_u4;
    row[0] = make_uint4(uint32_t(uint64_t(key)), uint32_t(uint64_t(key) >> 32), occ, pay[0]);
    if (jb.slot_u4 == 2) row[1] = make_uint4(pay[1], pay[2], pay[3], pay[4]);
  } else {
    atomicExch(P.table.overflow, 1u);
  }
  if (P.has_build_bloom) { bloom_insert(P.build_bloom, uint64_t(key)); ++n_bloom_ins; }
}

// ---- stage C: n entries on dense lanes: probe chains, second join, sink ---------------------
// (inlined at its single call site: an ABI call in the consumer loop makes ptxas keep the loop state on
// the stack)
template <uint32_t ACC>
__device__ __forceinline__ void stage_c(const DevPlan& P, const uint4* q, uint32_t n, uint32_t lane, uint32_t& n_out, uint32_t& n_bad,
                                        uint32_t& n_bloom_ins) {
  const bool act = lane < n;
  uint4 e = make_uint4(0, 0, 0, 0);
  if (act) e = q[lane];
  __syncwarp();
  GRow g;
  g.page = P.pages + uint64_t(e.x) * P.page_stride;
  PageDesc d{};
  if (act) d = P.descs[e.x];
  g.lc = P.classes + d.layout_class;
  g.r = e.y;
  g.nulls = d.null_mask & P.used_null_mask;
  g.pay0 = g.pay1 = nullptr;
  g.occ0 = g.occ1 = 0;
  JoinIter it0, it1;
  it0.cand = it1.cand = 0;
  it0.ended = it1.ended = true;
  it0.base = it0.tag = it0.klo = it0.khi = it1.base = it1.tag = it1.klo = it1.khi = 0;
  bool have0 = false;
  if (P.njoins) it0.init(P.joins[0], int64_t((uint64_t(e.w) << 32) | e.z), act);
  bool fresh = act;  // no join at all: every entry is emitted exactly once
  for (;;) {
    bool found = false;
    if (P.njoins == 0) {
      found = fresh;
      fresh = false;
    } else {
      for (;;) {
        if (have0) {
          if (P.njoins >= 2) {
            const uint32_t* s1 = it1.next(P.joins[1], &g.occ1);
            if (s1) { g.pay1 = s1; found = true; break; }
          }
          have0 = false;
        }
        const uint32_t* s0 = it0.next(P.joins[0], &g.occ0);
        if (!s0) break;
        g.pay0 = s0;
        have0 = true;
        if (P.njoins < 2) { found = true; break; }
        const DevJoin& j1 = P.joins[1];
        it1.init(j1, g_i64(j1.key, g), g_valid(j1.key, g));
      }
    }
    const uint32_t m = __ballot_sync(0xffffffffu, found);
    if (!m) break;
    n_out += __popc(m);
    if (P.sink == SINK_AGG) {
      if (P.nkeys) { if (found) sink_agg_grouped<ACC>(P, g, n_bad); }
      else sink_agg_single<ACC>(P, g, found, lane);
    } else if (P.sink == SINK_JOIN_BUILD) {
      sink_build(P, g, found, lane, n_bloom_ins);
    }
    __syncwarp();
  }
}

// FilterExec conjunct for one row of the staged tile (see term_pass2)
template <int LD, bool NONULL>
__device__ __forceinline__ bool term_pass1(const DevTerm& T, const uint8_t* stage, uint32_t r, uint32_t tile_nulls, uint32_t& bad) {
  const uint8_t* p = stage + T.ref.off;
  bool a;
  const uint32_t ld = LD >= 0 ? uint32_t(LD) : uint32_t(T.ref.ld);
  switch (ld) {
    case LD_F64: a = in_range1(f64_key(reinterpret_cast<const int64_t*>(p)[r]), T); break;
    case LD_VIEW: a = view_in_range(reinterpret_cast<const uint4*>(p)[r], T, bad); break;
    case LD_I32: a = in_range1(reinterpret_cast<const int32_t*>(p)[r], T); break;
    case LD_I64: a = in_range1(reinterpret_cast<const int64_t*>(p)[r], T); break;
    case LD_I16: a = in_range1(reinterpret_cast<const int16_t*>(p)[r], T); break;
    case LD_F32: {
      const int32_t x = reinterpret_cast<const int32_t*>(p)[r];
      a = in_range1(x ^ int32_t(uint32_t(x >> 31) >> 1), T);
      break;
    }
    default: {  // LD_DEC
      const uint4 v = reinterpret_cast<const uint4*>(p)[r];
      a = in_range2(((uint64_t(v.w) << 32) | v.z) ^ 0x8000000000000000ull, (uint64_t(v.y) << 32) | v.x, T);
      break;
    }
  }
  if (LD < 0 && T.op != TERM_IN_RANGE) a = T.op == TERM_NOT_IN_RANGE && !a;
  if (!NONULL && T.ref.valid_off != kNoValidity && ((tile_nulls >> T.ref.pcol) & 1))  // NULL => not TRUE => dropped
    a &= (stage[T.ref.valid_off + (r >> 3)] >> (r & 7)) & 1;
  return a;
}

// ---- the kernel ------------------------------------------------------------------------------
// T0 >= 0: the predicate is exactly one plain range term of load kind T0 over a NOT NULL column and
// no staged column is nullable (the Q3 pipelines); T0 < 0: generic.
template <uint32_t ACC, int T0>
__global__ void __launch_bounds__(kPThreads, 1) probe_pipeline_kernel(const __grid_constant__ DevPlan P) {
  constexpr bool kNoNull = T0 >= 0;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  ProbeShared* sh = reinterpret_cast<ProbeShared*>(smem_raw);
  uint8_t* stages = smem_raw + probe_shared_bytes();
  const uint32_t kNumStages = P.nstages;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < kPMaxStages; ++s) {
      mbar_init(&sh->full[s], 1);
      mbar_init(&sh->empty[s], kPConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  uint32_t n_in = 0, n_bloom = 0, n_filt = 0, n_out = 0, n_bad = 0, n_bloom_ins = 0;

  if (warp == 0) {
    // ===== producer (see pipeline_kernel.cuh): lanes 0..15 copy the values slice of staged column `lane`,
    // lanes 16..31 its validity slice; the page stream is marked evict_first in L2
    const uint64_t pol = l2_policy_evict_first();
    uint32_t ps = 0, pphase = 0;
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
      if (page + gridDim.x < P.npages) d_next = P.descs[page + gridDim.x];
      if (d.layout_class != cur_class) {
        cur_class = d.layout_class;
        const LayoutClass* lc = P.classes + cur_class;
        col_off = want ? (is_validity ? lc->validity_off[sc.page_col] : lc->values_off[sc.page_col]) : 0u;
      }
      const uint8_t* col_base = P.pages + uint64_t(page) * P.page_stride + col_off;
      const uint32_t null_mask = d.null_mask & P.used_null_mask;
      const bool active = want && (!is_validity || ((null_mask >> sc.page_col) & 1u));
      for (uint32_t tile = 0, r0 = 0; tile < P.tiles_per_page; ++tile, r0 += P.tile_rows) {
        const uint32_t s = ps;
        const uint32_t n = d.row_count > r0 ? min(d.row_count - r0, P.tile_rows) : 0u;
        uint32_t bytes = 0;
        if (active && n) bytes = is_validity ? ((((n + 7u) >> 3) + 15u) & ~15u) : ((n * width + 15u) & ~15u);
        const uint32_t total = __reduce_add_sync(0xffffffffu, bytes);
        mbar_wait(&sh->empty[s], pphase ^ 1u);
        if (++ps == kNumStages) { ps = 0; pphase ^= 1u; }
        if (lane == 0) {
          sh->meta[s] = PStageMeta{n, null_mask, page, r0};
          mbar_arrive_expect_tx(&sh->full[s], total);
        }
        __syncwarp();
        if (bytes)
          tma_load_1d_hint(stages + size_t(s) * P.stage_bytes + smem_off, col_base + (is_validity ? (r0 >> 3) : r0 * width), bytes, &sh->full[s], pol);
      }
    }
  } else {
    // ===== consumers =====
    // Stages A, B and C each appear exactly once, inlined, in one loop: per 32-row chunk stage A runs, then
    // stage B if 32 survivors are queued, then stage C if 32 tag hits are queued.  After the last tile the
    // same loop keeps turning with no chunk until both queues and the pending batch are drained.
    uint4* q1 = reinterpret_cast<uint4*>(stages + size_t(kNumStages) * P.stage_bytes) + size_t(warp - 1) * (2u * kPQueueEntries);
    uint4* q2 = q1 + kPQueueEntries;
    uint4* qa = P.njoins ? q1 : q2;             // without a join the survivors go straight to stage C
    uint32_t q1n = 0, q2n = 0;                  // warp-uniform fill levels
    const uint32_t lt = (1u << lane) - 1u;
    const uint64_t pol_keep = l2_policy_evict_last();
    // batch whose tag loads are in flight (stage B issued, not yet resolved)
    bool pend = false, pact = false;
    uint4 pe = make_uint4(0, 0, 0, 0);
    uint2 pw = make_uint2(0, 0);
    uint32_t ptag = 0;

    uint32_t cs = 0, cphase = 0, cb = 0;  // ring stage / phase; chunks dealt so far modulo the warp count
    const uint32_t my_pages = P.npages > blockIdx.x ? (P.npages - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
    const uint32_t my_items = my_pages * P.tiles_per_page;
    for (uint32_t item = 0; item <= my_items; ++item) {
      const bool last = item == my_items;  // the drain turn
      uint32_t s = 0, nchunks = 0, c = 0;
      PStageMeta meta{0, 0, 0, 0};
      const uint8_t* stage = stages;
      if (!last) {
        s = cs;
        mbar_wait(&sh->full[s], cphase);
        if (++cs == kNumStages) { cs = 0; cphase ^= 1u; }
        meta = sh->meta[s];
        stage = stages + size_t(s) * P.stage_bytes;
        nchunks = (meta.nrows + 31u) >> 5;
        // 32-row chunks are dealt to the warps round-robin ACROSS tiles, so the load stays balanced
        // whatever the tile size; the ring lets a warp run ahead of the others by its depth
        c = warp - 1u >= cb ? warp - 1u - cb : warp - 1u + kPConsumerWarps - cb;
        cb = (cb + nchunks) % kPConsumerWarps;
      }
      for (;;) {
        const bool have = c < nchunks;
        if (!have && !(last && (q1n | q2n | uint32_t(pend)))) break;
        if (have) {
          // -- stage A: runtime Bloom probes (NULL key => DefinitelyAbsent, shared.rs:367-374), conjuncts,
          // compaction of the survivors
          const uint32_t r = c * 32u + lane;
          const bool has = r < meta.nrows;
          const uint32_t rr = has ? r : 0u;  // row 0 of a tile always exists
          const uint32_t nhere = min(32u, meta.nrows - c * 32u);
          c += kPConsumerWarps;
          bool keep = has;
          n_in += nhere;
          for (uint32_t b = 0; b < P.nbloom; ++b) {
            const DevBloomProbe& bp = P.bloom[b];
            const Row rq{stage, rr, meta.null_mask, nullptr, 0};
            bool k1[1] = {keep && ref_valid(bp.key, rq)};
            const uint64_t bk[1] = {uint64_t(load_i64(bp.key, rq))};
            bloom_contains_n<1>(bp.bloom, bk, k1);
            keep = k1[0];
          }
          if (P.nbloom) n_bloom += __popc(__ballot_sync(0xffffffffu, keep));
          else n_bloom += nhere;
          if constexpr (T0 >= 0) {
            keep = keep && term_pass1<T0, true>(P.terms[0], stage, rr, 0u, n_bad);
          } else {
            for (uint32_t t = 0; t < P.nterms; ++t) {
              if (!__any_sync(0xffffffffu, keep)) break;
              keep = keep && term_pass1<-1, false>(P.terms[t], stage, rr, meta.null_mask, n_bad);
            }
          }
          uint32_t m = __ballot_sync(0xffffffffu, keep);
          n_filt += __popc(m);
          int64_t key = 0;
          if (P.njoins) {
            const DevJoin& j = P.joins[0];
            const Row rq{stage, rr, meta.null_mask, nullptr, 0};
            key = load_i64(j.key, rq);
            if constexpr (!kNoNull) {
              keep = keep && ref_valid(j.key, rq);  // NULL keys never match: an inner join drops the row
              m = __ballot_sync(0xffffffffu, keep);
            }
          }
          if (m) {
            const uint32_t qn = P.njoins ? q1n : q2n;
            if (keep) qa[qn + __popc(m & lt)] = make_uint4(meta.page, meta.r0 + rr, uint32_t(uint64_t(key)), uint32_t(uint64_t(key) >> 32));
            if (P.njoins) q1n += __popc(m);
            else q2n += __popc(m);
            __syncwarp();
          }
        }
        // -- stage B: resolve the tag loads of the pending batch (SIMD-in-register byte compares) and compact
        // the hits into q2; then pop the next batch of survivors onto dense lanes, hash, issue its tag loads
        if (q1n >= 32u || (!have && (q1n | uint32_t(pend)))) {
          if (pend) {
            pend = false;
            const TagWindow tw = scan_tags(pw, ptag);
            const bool hit = pact && (tw.cand != 0ull || !tw.ended);
            const uint32_t m = __ballot_sync(0xffffffffu, hit);
            if (m) {
              if (hit) q2[q2n + __popc(m & lt)] = pe;
              q2n += __popc(m);
              __syncwarp();
            }
          }
          const uint32_t n = min(q1n, 32u);
          if (n) {
            q1n -= n;
            pact = lane < n;
            if (pact) pe = q1[q1n + lane];
            __syncwarp();
            const DevJoin& j = P.joins[0];
            const uint64_t h = join_hash(int64_t((uint64_t(pe.w) << 32) | pe.z));
            ptag = join_tag8(h, j.shift);
            pw = ldg_tags8(j.tags + join_home(h, j.shift), pact, pol_keep);
            pend = true;
          }
        }
        // -- stage C: 32 tag hits (all that is left on the drain turn) on dense lanes
        if (q2n >= 32u || (!have && q2n && !pend && !q1n)) {
          const uint32_t n = min(q2n, 32u);
          q2n -= n;
          stage_c<ACC>(P, q2 + q2n, n, lane, n_out, n_bad, n_bloom_ins);
        }
      }
      if (!last) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh->empty[s]);
      }
    }
  }

  // counters (RuntimeFilter*/Worker* style metrics).  n_in .. n_out are warp-uniform in the consumer
  // warps (ballot popcounts): one lane adds them; n_bad and n_bloom_ins are per lane.
  {
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(P.counters);
    if (lane == 0 && warp > 0) {
      if (n_in) atomicAdd(dst + 0, (unsigned long long)n_in);
      if (n_bloom) atomicAdd(dst + 1, (unsigned long long)n_bloom);
      if (n_filt) atomicAdd(dst + 2, (unsigned long long)n_filt);
      if (n_out) atomicAdd(dst + 3, (unsigned long long)n_out);
    }
    uint32_t vals[2] = {n_bloom_ins, n_bad};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      uint32_t v = vals[q];
#pragma unroll
      for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && v) atomicAdd(dst + 4 + q, (unsigned long long)v);
    }
  }
}

}  // namespace pgf
