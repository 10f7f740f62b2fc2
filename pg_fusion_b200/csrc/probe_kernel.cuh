// The fused scan pipeline for selective plans: anything with a HashJoinExec probe or a
// HashJoinExec build sink (sm_100a).
//
// One persistent CTA per SM, and inside it every warp is its own pipeline: it owns a private ring of
// row tiles in shared memory which it fills itself with TMA bulk copies (cp.async.bulk + mbarrier
// transaction counts, marked evict_first in L2) a few tiles ahead of where it reads.  There is no
// producer warp and no CTA-wide barrier, so a warp that is busy in the rare, latency-heavy stage C
// below never holds up the others.  The work of a warp is organised around *selection-vector
// compaction* instead of per-thread predication:
//
//   stage A  filter      32 rows at a time from the staged tile: runtime Bloom probes and the FilterExec
//                        conjuncts in registers, then ballot + prefix-popcount compaction of the surviving
//                        (page, row, join key) triples into a per-warp queue in shared memory.  Only the
//                        predicate columns and the probe key are staged at all.
//   stage B  tag probe   whenever the queue holds 32 survivors they are popped onto DENSE lanes, hashed,
//                        and one aligned 8-byte load of the home bucket's tag bytes is issued per lane (L2
//                        resident directory, evict_last).  The loads stay in flight while the warp goes
//                        back to stage A; they are resolved (SIMD-in-register byte compares) just before
//                        the next batch is issued.  Tag hits are compacted into a second queue.
//   stage C  match+sink  32 tag hits at a time, again on dense lanes: walk the probe chain, compare the
//                        slot keys (duplicates multiply, NULL keys never match), optionally probe a
//                        second join table, and feed the sink.  Columns that only matched rows need
//                        (aggregate arguments, group keys, build payloads) are read straight from the
//                        page in HBM here -- late materialisation: for TPC-H Q3's lineitem side that is
//                        1 % of the rows, so 16 of the 36 algorithmic bytes per row never leave DRAM.
//
// Sinks: AggregateExec into the global group table (rows of one group inside a warp elect a leader with
// match.any, so a group is looked up once per warp and batch) or warp-reduced into slot 0 (no GROUP BY);
// HashJoinExec build side as a dense array of build rows (warp-aggregated append; the table is built
// from the rows afterwards at exactly the capacity the row count asks for, see join_build_kernel in
// pipeline.cu) [+ RuntimeFilterBuildExec]; row count.
#pragma once
#include "pipeline_kernel.cuh"

namespace pgf {

#ifndef PGF_PROBE_WARPS
#define PGF_PROBE_WARPS 20
#endif
constexpr int kPConsumerWarps = PGF_PROBE_WARPS;
constexpr int kPThreads = kPConsumerWarps * 32;
constexpr uint32_t kPMaxDepth = 4;                                   // tiles in flight per warp
constexpr uint32_t kPQueueEntries = 64;                              // per warp and queue; drained 32 at a time
constexpr uint32_t kPQueueBytesPerWarp = 2u * kPQueueEntries * 16u + 32u * 8u;  // survivors + tag hits + the tag windows in flight

struct PStageMeta {
  uint32_t nrows, null_mask, page, r0;
};
struct PWarpCtl {  // one per warp
  uint64_t full[kPMaxDepth];
  PStageMeta meta[kPMaxDepth];
  // counters that are touched once per tile or only by stages B / C: kept here, not in registers of the stage A loop
  uint32_t n_in, n_out, n_bad, n_groups, n_bloom_rej, pad0;
  uint64_t dbar;       // completes when the descriptor of the next tile's page has landed in dnext
  PageDesc dnext;
  uint8_t pad1[16];
};
static_assert(sizeof(PWarpCtl) == 160 && offsetof(PWarpCtl, dnext) % 16 == 0, "per-warp control block");
__host__ __device__ constexpr uint32_t probe_shared_bytes() { return uint32_t(kPConsumerWarps) * uint32_t(sizeof(PWarpCtl)); }

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
// The same window as an asynchronous copy into the lane's shared-memory slot.  A register destination would tie
// the load to the scoreboard every other global load of the kernel shares, and the compiler waits on that one at
// the top of the next chunk; the copy is waited for where stage B resolves the batch, a whole batch later.
__device__ __forceinline__ void tags8_async(uint32_t smem_dst, const uint8_t* p, bool pred, uint64_t policy) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 8, %3;\n\t}"
               ::"r"(smem_dst), "l"(p), "r"(uint32_t(pred)), "l"(policy) : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
}
#ifndef PGF_LATE_PREFETCH
#define PGF_LATE_PREFETCH 1   // 0: off, 1: into L1, 2: into L2
#endif
__device__ __forceinline__ void prefetch_line(const void* p) {
#if PGF_LATE_PREFETCH == 1
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#elif PGF_LATE_PREFETCH == 2
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}
__device__ __forceinline__ uint2 tags8_collect(uint32_t smem_src) {
  uint2 v;
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(smem_src) : "memory");
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

// Stage B only needs to know whether the window is worth a visit by stage C: some tag before the first empty one
// equals `tag`, or the window holds no empty tag at all (the chain goes on in the next bucket).
__device__ __forceinline__ bool window_may_match(uint2 w, uint32_t tag) {
  const uint32_t t4 = tag * 0x01010101u;
  const uint32_t z0 = zero_bytes(w.x), z1 = zero_bytes(w.y), m0 = zero_bytes(w.x ^ t4), m1 = zero_bytes(w.y ^ t4);
  const uint32_t before0 = z0 ? (z0 & (0u - z0)) - 1u : 0xFFFFFFFFu;   // bytes of the low word before its first empty tag
  const uint32_t before1 = z0 ? 0u : (z1 ? (z1 & (0u - z1)) - 1u : 0xFFFFFFFFu);
  return ((m0 & before0) | (m1 & before1) | uint32_t((z0 | z1) == 0u)) != 0u;
}

// ---- a row of a page in HBM (stage C reads what it needs from there) ----------------------
struct GRow {
  const uint8_t* page;
  const LayoutClass* lc;
  uint32_t r;
  uint32_t nulls;                  // page columns with nulls in this page (already masked by use)
  const uint32_t *pay0, *pay1;     // matched join slots (u32 words: key lo, key hi, occ, payload...)
  const uint32_t* rec;             // row-set scans: the scanned row itself, same record format (source kSrcRecord)
  uint32_t occ0, occ1, occr;
  __device__ __forceinline__ const uint32_t* pay(uint32_t src) const { return src == 1 ? pay0 : (src == 2 ? pay1 : rec); }
  __device__ __forceinline__ uint32_t occ(uint32_t src) const { return src == 1 ? occ0 : (src == 2 ? occ1 : occr); }
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
  const uint32_t* p = (ref.src == kSrcRecord ? g.rec + ref.off : g.pay(ref.src) + 3 + ref.off);
  switch (ref.ld) {
    case LD_I16: return int64_t(int16_t(__ldg(p)));
    case LD_I32: return int64_t(int32_t(__ldg(p)));
    default: return int64_t((uint64_t(__ldg(p + 1)) << 32) | __ldg(p));
  }
}
__device__ __forceinline__ uint4 g_u128(const DevRef& ref, const GRow& g) {
  if (ref.src == SRC_PAGE) return reinterpret_cast<const uint4*>(g.page + g.lc->values_off[ref.pcol])[g.r];
  const uint32_t* p = (ref.src == kSrcRecord ? g.rec + ref.off : g.pay(ref.src) + 3 + ref.off);
  return make_uint4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
}
__device__ __forceinline__ double g_f64(const DevRef& ref, const GRow& g) {
  if (ref.src == SRC_PAGE) {
    const uint8_t* p = g.page + g.lc->values_off[ref.pcol];
    if (ref.ld == LD_F64) return reinterpret_cast<const double*>(p)[g.r];
    if (ref.ld == LD_F32) return double(reinterpret_cast<const float*>(p)[g.r]);
    return double(g_i64(ref, g));  // AVG over integers runs on the Float64 cast
  }
  const uint32_t* p = (ref.src == kSrcRecord ? g.rec + ref.off : g.pay(ref.src) + 3 + ref.off);
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
    // (with the evict_last policy stages A / B give the directory: a plain load here made the lines ordinary again)
    const uint2 w = ldg_tags8(j.tags + base, true, l2_policy_evict_last());
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
  // next slot of the chain whose TAG equals the probe key's (its key is still to be compared), or nullptr
  __device__ __forceinline__ const uint4* next_candidate(const DevJoin& j) {
    for (;;) {
      if (cand) {
        const uint32_t b = uint32_t(__ffsll((long long)cand)) - 1u;
        cand &= cand - 1ull;
        return j.slots + uint64_t(base + (b >> 3)) * j.slot_u4;
      }
      if (ended) return nullptr;
      base = (base + kJoinBucket) & j.mask;
      load(j);
    }
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
// group_slot() without waiting: -2 when the probe runs into a slot another thread is publishing.  The
// caller retries from a converged point, so lanes of one warp never spin on each other.
// (*inserted is set when the call created the group: the caller counts new groups per thread and adds them to
// GroupTable::used once per warp at the end of the kernel -- one same-address atomic per group serialises in L2.)
__device__ __forceinline__ int64_t group_slot_try(const GroupTable& t, const uint64_t* key, uint32_t nwords, uint32_t knull, uint64_t h,
                                                  uint32_t* inserted) {
  uint32_t i = uint32_t(h) & t.mask;
  const uint32_t ready = 2u | (knull << 8);
  for (uint32_t probes = 0; probes <= t.mask; ++probes, i = (i + 1) & t.mask) {
    uint32_t s = ld_acquire_u32(t.state + i);
    if (s == 0) {
      const uint32_t old = atomicCAS(t.state + i, 0u, 1u);
      if (old == 0) {
        group_slot_init(t, i, key, nwords);
        __threadfence();
        atomicExch(t.state + i, ready);
        *inserted += 1u;
        return i;
      }
      s = old == 1u ? old : ld_acquire_u32(t.state + i);   // (the CAS itself is relaxed)
    }
    if ((s & 3u) == 1u) return -2;
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

// Called by all 32 lanes; `found` lanes carry a row.  Rows of one group that meet in a batch (the
// lineitems of an order are neighbours) elect a leader with match.any: one table lookup per group, warp
// and batch, and no two lanes of a warp ever insert the same key.
template <uint32_t ACC>
__device__ __forceinline__ void sink_agg_grouped(const DevPlan& P, const GRow& g, bool found, uint32_t lane, uint32_t& bad, uint32_t& new_groups) {
  using Ops = AccOps<ACC>;
  uint64_t key[kKeyWords] = {0, 0, 0, 0};
  uint32_t knull = 0;
  if (found) {
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
  }
  const uint64_t h = found ? key_hash(key, P.nkeywords, knull) : (0xFFFFFFFF00000000ull | lane);
  const uint32_t peers = __match_any_sync(0xffffffffu, h);
  const uint32_t leader = uint32_t(__ffs(int(peers))) - 1u;
  bool same = __shfl_sync(0xffffffffu, knull, leader) == knull;
#pragma unroll
  for (uint32_t w = 0; w < kKeyWords; ++w) same &= __shfl_sync(0xffffffffu, key[w], leader) == key[w];
  const bool own = found && (lane == leader || !same);  // (a hash collision inside the warp: that lane looks up its own key)
  int64_t slot = -1;
  bool todo = own;
  while (__any_sync(0xffffffffu, todo)) {
    if (todo) {
      const int64_t s = group_slot_try(P.table, key, P.nkeywords, knull, h, &new_groups);
      if (s != -2) { slot = s; todo = false; }
    }
  }
  const int64_t lslot = __shfl_sync(0xffffffffu, slot, leader);
  if (found && !own) slot = lslot;
  if (!found || slot < 0) return;
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
// (RuntimeFilterBuildExec: the filter is populated from the dense rows afterwards, bloom_insert_rows_kernel in pipeline.cu)
__device__ __forceinline__ void sink_build(const DevPlan& P, const GRow& g, bool found, uint32_t lane) {
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
}

// ---- stage C: n entries on dense lanes: probe chains, second join, sink ---------------------
// (inlined at its single call site: an ABI call in the consumer loop makes ptxas keep the loop state on
// the stack)
// e: {page, row, key lo, key hi} of a page scan, or {row index lo, hi, key lo, key hi} of a row-set scan (ROWS).
template <uint32_t ACC, bool ROWS = false>
__device__ __forceinline__ void stage_c(const DevPlan& P, uint4 e, bool act, uint32_t lane, uint32_t& n_out, uint32_t& n_bad,
                                        uint32_t& n_groups) {
  GRow g;
  g.pay0 = g.pay1 = g.rec = nullptr;
  g.occ0 = g.occ1 = g.occr = 0;
  if constexpr (ROWS) {
    g.page = nullptr;
    g.lc = nullptr;
    g.r = 0;
    g.nulls = 0;
    g.rec = reinterpret_cast<const uint32_t*>(P.row_src + ((uint64_t(e.y) << 32) | e.x) * P.row_u4);
    if (act) g.occr = __ldg(g.rec + 2);
  } else {
    g.page = P.pages + uint64_t(e.x) * P.page_stride;
    g.r = e.y;
    if (P.single_class && !P.used_null_mask) {   // the common case: offsets from the constant bank, no descriptor load
      g.lc = &P.class0;
      g.nulls = 0;
    } else {
      PageDesc d{};
      if (act) d = P.descs[e.x];
      g.lc = P.single_class ? &P.class0 : P.classes + d.layout_class;
      g.nulls = d.null_mask & P.used_null_mask;
    }
    // Build sinks copy every late column of every row that finds its partner, and behind a selective predicate most
    // entries that reach stage C do: start the reads of the row's late columns now, so that they are in flight while
    // the chain walk waits for the tag window and the slot (ncu r2: 40 % of the stall samples of the orders pipeline
    // sat on these loads, one HBM round trip after the other; -6 % on that pipeline).  Not for aggregate sinks: two
    // thirds of the lineitem side's tag hits are false, and the prefetch cost it 8 % at SF100.
    if (PGF_LATE_PREFETCH && P.sink == SINK_JOIN_BUILD && act)
      for (uint32_t c = 0; c < P.nlate; ++c) prefetch_line(g.page + g.lc->values_off[P.late_pcol[c]] + g.r * uint32_t(P.late_width[c]));
  }
  JoinIter it0, it1;
  it0.cand = it1.cand = 0;
  it0.ended = it1.ended = true;
  it0.base = it0.tag = it0.klo = it0.khi = it1.base = it1.tag = it1.klo = it1.khi = 0;
  bool have0 = false;
  if (P.njoins) {
    int64_t k0 = int64_t((uint64_t(e.w) << 32) | e.z);
    bool v0 = act;
    if constexpr (ROWS) {  // any column of the scanned row may be the probe key
      if (act) { k0 = g_i64(P.joins[0].key, g); v0 = g_valid(P.joins[0].key, g); }
    }
    it0.init(P.joins[0], k0, v0);
  }
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
      if (P.nkeys) sink_agg_grouped<ACC>(P, g, found, lane, n_bad, n_groups);
      else sink_agg_single<ACC>(P, g, found, lane);
    } else if (P.sink == SINK_JOIN_BUILD) {
      sink_build(P, g, found, lane);
    }
    __syncwarp();
  }
}

// Inline view in range, the way view_in_range() defines it -- the key is the 128-bit number (big-endian bytes 0..3,
// 4..7, 8..11, length) and the test is (key - lo) <= span, unsigned -- with explicit 32-bit carry chains: 3 byte
// permutes + 9 integer instructions, against ~45 for the compiler's generic 128-bit arithmetic.
__device__ __forceinline__ bool view_in_range_fast(uint4 v, const DevTerm& T, const DevPlan& P, uint32_t page) {
  v = view_first12(v, P, page);   // out-of-line value: its bytes 4..11 come from the page's tail arena
  const uint32_t w0 = bswap32(v.y), w1 = bswap32(v.z), w2 = bswap32(v.w), w3 = v.x;
  const uint32_t l0 = uint32_t(uint64_t(T.lo0) >> 32), l1 = uint32_t(uint64_t(T.lo0)), l2 = uint32_t(T.lo1 >> 32), l3 = uint32_t(T.lo1);
  const uint32_t s0 = uint32_t(uint64_t(T.hi0) >> 32), s1 = uint32_t(uint64_t(T.hi0)), s2 = uint32_t(T.hi1 >> 32), s3 = uint32_t(T.hi1);
  uint32_t borrow;
  asm("{\n\t.reg .u32 d0, d1, d2, d3, t;\n\t"
      "sub.cc.u32 d3, %1, %5;\n\tsubc.cc.u32 d2, %2, %6;\n\tsubc.cc.u32 d1, %3, %7;\n\tsubc.u32 d0, %4, %8;\n\t"       // d = key - lo
      "sub.cc.u32 t, %9, d3;\n\tsubc.cc.u32 t, %10, d2;\n\tsubc.cc.u32 t, %11, d1;\n\tsubc.cc.u32 t, %12, d0;\n\t"   // span - d
      "subc.u32 %0, 0, 0;\n\t}"                                                                                    // 0 or 0xFFFFFFFF = borrow
      : "=r"(borrow)
      : "r"(w3), "r"(w2), "r"(w1), "r"(w0), "r"(l3), "r"(l2), "r"(l1), "r"(l0), "r"(s3), "r"(s2), "r"(s1), "r"(s0));
  return borrow == 0u;
}

// FilterExec conjunct for one row of the staged tile (see term_pass2)
template <int LD, bool NONULL>
__device__ __forceinline__ bool term_pass1(const DevTerm& T, const uint8_t* stage, uint32_t r, uint32_t tile_nulls, const DevPlan& P, uint32_t page) {
  const uint8_t* p = stage + T.ref.off;
  bool a;
  const uint32_t ld = LD >= 0 ? uint32_t(LD) : uint32_t(T.ref.ld);
  switch (ld) {
    case LD_F64: a = in_range1(f64_key(reinterpret_cast<const int64_t*>(p)[r]), T); break;
    case LD_VIEW: a = view_in_range_fast(reinterpret_cast<const uint4*>(p)[r], T, P, page); break;
    case LD_I32: a = in_range1(reinterpret_cast<const int32_t*>(p)[r], T); break;
    case LD_I64: a = in_range1(reinterpret_cast<const int64_t*>(p)[r], T); break;
    case LD_I16: a = in_range1(reinterpret_cast<const int16_t*>(p)[r], T); break;
    case LD_F32: {
      const int32_t x = reinterpret_cast<const int32_t*>(p)[r];
      a = in_range1(x ^ int32_t(uint32_t(x >> 31) >> 1), T);
      break;
    }
    case LD_BOOL: a = in_range1(int64_t((p[r >> 3] >> (r & 7u)) & 1u), T); break;   // bit-packed values
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
// T0 >= 0 (the Q3 pipelines): the predicate is exactly one plain range term of load kind T0 over a NOT NULL
// column, no staged column is nullable, every runtime filter is probed on dense lanes (stage B) and the entry key
// (join key or dense filter key) is an Int32 page column staged at P.entry_key_off; T0 < 0: generic.
// SPLIT: stage C is not part of this kernel -- the tag hits are appended to P.entries and entries_pipeline_kernel
// below runs stage C over them.  Stage C is rare but slow: a batch walks dependent HBM round trips (slot, late
// columns, group table) and ~1500 divergent instructions while the warp's tile stream stands still; with it compiled
// out the same 20 warps scan SF100's lineitem side in 2.6 instead of 4.8 ms, and its 96 registers become 64.
template <uint32_t ACC, int T0, bool SPLIT = false>
__global__ void __launch_bounds__(kPThreads, 1) probe_pipeline_kernel(const __grid_constant__ DevPlan P) {
  constexpr bool kNoNull = T0 >= 0;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t D = P.nstages;  // tiles in flight per warp
  // per-warp shared-memory areas.  The offsets are made opaque to the compiler once, so that it keeps them in
  // registers instead of re-deriving them from threadIdx and the plan inside the hot loop (7 % of the instructions
  // of the first version of this kernel).
  uint32_t ctl_off = warp * uint32_t(sizeof(PWarpCtl));
  uint32_t stages_off = probe_shared_bytes() + warp * D * P.stage_bytes;
  uint32_t q_off = probe_shared_bytes() + uint32_t(kPConsumerWarps) * D * P.stage_bytes + warp * kPQueueBytesPerWarp;
  asm volatile("" : "+r"(ctl_off), "+r"(stages_off), "+r"(q_off));
  PWarpCtl* ctl = reinterpret_cast<PWarpCtl*>(smem_raw + ctl_off);
  uint8_t* stages = smem_raw + stages_off;
  uint4* q1 = reinterpret_cast<uint4*>(smem_raw + q_off);
  uint4* q2 = q1 + kPQueueEntries;
  const uint32_t tagslot = smem_u32(q2 + kPQueueEntries) + lane * 8u;   // this lane's tag window in flight
  if (lane == 0) {
    ctl->n_in = ctl->n_out = ctl->n_bad = ctl->n_groups = ctl->n_bloom_rej = 0;
    mbar_init(&ctl->dbar, 1);
    for (uint32_t s = 0; s < kPMaxDepth; ++s) mbar_init(&ctl->full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();

  uint32_t n_bloom = 0, n_filt = 0, n_bad = 0;   // (n_bad: stage A's out-of-line views; the rest of the counters: PWarpCtl)

  // ---- this warp's tile stream: tiles wg, wg + nw, wg + 2 nw, ... of the scan (tile = tile_rows rows of a page)
  const uint32_t wg = blockIdx.x * kPConsumerWarps + warp, nw = gridDim.x * kPConsumerWarps;
  const uint64_t total_tiles = uint64_t(P.npages) * P.tiles_per_page;
  const uint32_t my_tiles = total_tiles > wg ? uint32_t((total_tiles - wg + nw - 1) / nw) : 0u;
  uint32_t ipage = wg / P.tiles_per_page, itip = wg % P.tiles_per_page;  // issue cursor
  const uint32_t dq = nw / P.tiles_per_page, dr = nw % P.tiles_per_page;
  uint32_t issued = 0, is = 0;
  // lanes 0..15 copy the values slice of staged column `lane`, lanes 16..31 its validity slice
  const uint64_t pol_stream = l2_policy_evict_first();
  const uint32_t mycol = lane & 15u;
  const bool is_validity = lane >= 16;
  bool want = false;
  uint32_t my_width = 0, my_soff = 0, my_pcol = 0, my_coff = 0, cur_class = 0xFFFFFFFFu;
  if (mycol < P.nstage_cols) {
    const DevStageCol sc = P.scol[mycol];
    want = !is_validity || sc.nullable;
    my_width = is_validity ? 0u : uint32_t(sc.width);
    my_soff = is_validity ? sc.valid_off : sc.smem_off;
    my_pcol = sc.page_col;
  }
  // The descriptor of the page of the next tile to issue is fetched one tile ahead by a bulk copy of its own (an
  // ordinary load would be waited for at the top of the next chunk, see tags8_async): parity of fetch i is i & 1.
  auto fetch_desc = [&]() {
    if (lane == 0) {
      mbar_arrive_expect_tx(&ctl->dbar, uint32_t(sizeof(PageDesc)));
      tma_load_1d_hint(&ctl->dnext, P.descs + ipage, uint32_t(sizeof(PageDesc)), &ctl->dbar, pol_stream);
    }
  };
  if (my_tiles) fetch_desc();
  auto issue_next = [&]() {
    if (issued >= my_tiles) return;
    mbar_wait(&ctl->dbar, issued & 1u);
    const uint2 dw = *reinterpret_cast<const uint2*>(&ctl->dnext);   // {row_count, layout_class | null_mask << 16}
    struct { uint32_t row_count, layout_class, null_mask; } d{dw.x, dw.y & 0xFFFFu, dw.y >> 16};
    if (d.layout_class != cur_class) {  // rare: pages of one scan share their layout class
      cur_class = d.layout_class;
      const LayoutClass* lc = P.classes + cur_class;
      my_coff = want ? (is_validity ? lc->validity_off[my_pcol] : lc->values_off[my_pcol]) : 0u;
    }
    const uint32_t r0 = itip * P.tile_rows;
    const uint32_t n = d.row_count > r0 ? min(d.row_count - r0, P.tile_rows) : 0u;
    const uint32_t null_mask = d.null_mask & P.used_null_mask;
    const bool active = want && n && (!is_validity || ((null_mask >> my_pcol) & 1u));
    uint32_t bytes = 0;
    if (active) bytes = my_width == 0u ? ((((n + 7u) >> 3) + 15u) & ~15u) : ((n * my_width + 15u) & ~15u);   // (width 0: a bitmap -- validity, or Boolean values)
    const uint32_t total = __reduce_add_sync(0xffffffffu, bytes);
    if (lane == 0) {
      ctl->meta[is] = PStageMeta{n, null_mask, ipage, r0};
      ctl->n_in += n;
      mbar_arrive_expect_tx(&ctl->full[is], total);
    }
    __syncwarp();   // (also: every lane has read dnext before the next fetch overwrites it)
    if (bytes)
      tma_load_1d_hint(stages + size_t(is) * P.stage_bytes + my_soff,
                       P.pages + uint64_t(ipage) * P.page_stride + my_coff + (my_width == 0u ? (r0 >> 3) : r0 * my_width), bytes, &ctl->full[is], pol_stream);
    ipage += dq;
    itip += dr;
    if (itip >= P.tiles_per_page) { itip -= P.tiles_per_page; ++ipage; }
    ++issued;
    if (++is == D) is = 0;
    if (issued < my_tiles) fetch_desc();  // in flight while the warp works on its next tile
  };
  for (uint32_t i = 0; i < D; ++i) issue_next();

  // Per 32-row chunk: stage A, then stage B once 32 survivors are queued, then stage C once 32 tag hits are
  // queued.  The hot loop carries no drain logic; what is left in the queues after the last tile is drained by
  // the same (out-of-line copies of the) stages below.
  {
    uint32_t q1n = 0, q2n = 0;                  // warp-uniform fill levels
    const uint32_t lt = (1u << lane) - 1u;
    const uint64_t pol_keep = l2_policy_evict_last();
    // batch whose tag loads are in flight (stage B issued, not yet resolved)
    bool pend = false, pact = false;
    uint4 pe = make_uint4(0, 0, 0, 0);
    uint32_t ptag = 0;

    // -- stage C: n tag hits (32, or what is left when draining) on dense lanes
    auto run_c = [&](uint32_t n) {
      q2n -= n;
      uint4 e = make_uint4(0, 0, 0, 0);
      if (lane < n) e = q2[q2n + lane];
      __syncwarp();
      if constexpr (SPLIT) {   // hand the batch to the stage-C kernel: one reservation per warp, one coalesced store
        unsigned long long at = 0;
        if (lane == 0) at = atomicAdd(P.entries_count, (unsigned long long)n);
        at = __shfl_sync(0xffffffffu, at, 0) + lane;
        if (lane < n) {
          if (at < P.entries_cap) P.entries[at] = e;
          else atomicExch(P.entries_overflow, 1u);
        }
        return;
      }
      uint32_t c_out = 0, c_bad = 0, c_groups = 0;
      stage_c<ACC>(P, e, lane < n, lane, c_out, c_bad, c_groups);
      if (lane == 0) ctl->n_out += c_out;               // warp uniform
      if (c_bad) atomicAdd(&ctl->n_bad, c_bad);         // per lane, rare
      if (c_groups) atomicAdd(&ctl->n_groups, c_groups);
    };
    // -- stage B: resolve the tag loads of the pending batch (SIMD-in-register byte compares), compact its hits
    // into q2; then pop n survivors onto dense lanes, hash them and issue their tag loads
    auto run_b = [&](uint32_t n) {
      if (pend) {
        pend = false;
        const bool hit = pact && window_may_match(tags8_collect(tagslot), ptag);
        const uint32_t m = __ballot_sync(0xffffffffu, hit);
        if (m) {
          if (hit) q2[q2n + __popc(m & lt)] = pe;
          q2n += __popc(m);
          __syncwarp();
        }
      }
      if (n) {
        q1n -= n;
        pact = lane < n;
        if (pact) pe = q1[q1n + lane];
        __syncwarp();
        if (P.bloom_dense) {  // the runtime filter on the entry key, probed on dense lanes
          const bool alive = pact && bloom_contains(P.bloom[0].bloom, (uint64_t(pe.w) << 32) | pe.z);
          const uint32_t rej = __popc(__ballot_sync(0xffffffffu, pact && !alive));
          if (lane == 0) ctl->n_bloom_rej += rej;
          pact = alive;
        }
        if (P.njoins) {
          const DevJoin& j = P.joins[0];
          const uint64_t h = join_hash(int64_t((uint64_t(pe.w) << 32) | pe.z));
          ptag = join_tag8(h, j.shift);
          tags8_async(tagslot, j.tags + join_home(h, j.shift), pact, pol_keep);
          pend = true;
        } else {   // no join: the survivors are the sink's rows
          const uint32_t m = __ballot_sync(0xffffffffu, pact);
          if (pact) q2[q2n + __popc(m & lt)] = pe;
          q2n += __popc(m);
          __syncwarp();
        }
      }
    };

    uint32_t cs = 0, cphase = 0;  // ring stage / mbarrier phase of the tile being read
    for (uint32_t k = 0; k < my_tiles; ++k) {
      mbar_wait(&ctl->full[cs], cphase);
      const PStageMeta meta = ctl->meta[cs];
      const uint8_t* stage = stages + size_t(cs) * P.stage_bytes;
      if (++cs == D) { cs = 0; cphase ^= 1u; }
      for (uint32_t r0 = 0; r0 < meta.nrows; r0 += 32u) {
        // -- stage A: runtime Bloom probes (NULL key => DefinitelyAbsent, shared.rs:367-374), conjuncts,
        // compaction of the survivors
        const uint32_t r = r0 + lane;
        const bool has = r < meta.nrows;
        const uint32_t rr = has ? r : 0u;  // row 0 of a tile always exists
        bool keep = has;
        if (T0 < 0 && P.nbloom > P.bloom_dense) {
          for (uint32_t b = P.bloom_dense; b < P.nbloom; ++b) {
            const DevBloomProbe& bp = P.bloom[b];
            const Row rq{stage, rr, meta.null_mask, nullptr, 0};
            bool k1[1] = {keep && ref_valid(bp.key, rq)};
            const uint64_t bk[1] = {uint64_t(load_i64(bp.key, rq))};
            bloom_contains_n<1>(bp.bloom, bk, k1);
            keep = k1[0];
          }
          n_bloom += __popc(__ballot_sync(0xffffffffu, keep));
        }
        // the entry's key (join key, or the key of the dense runtime filter): its load is issued ahead of the
        // conjuncts', so that one wait covers both
        int64_t key = 0;
        bool kvalid = true;
        if constexpr (T0 >= 0) {
          key = reinterpret_cast<const int32_t*>(stage + P.entry_key_off)[rr];
        } else if (P.njoins | P.bloom_dense) {
          const DevRef& kr = P.njoins ? P.joins[0].key : P.bloom[0].key;
          const Row rq{stage, rr, meta.null_mask, nullptr, 0};
          key = load_i64(kr, rq);
          if constexpr (!kNoNull) kvalid = ref_valid(kr, rq);
        }
        if constexpr (T0 >= 0) {
          keep = keep & term_pass1<T0, true>(P.terms[0], stage, rr, 0u, P, meta.page);   // (row rr exists on every lane: no branch)
        } else {
          for (uint32_t t = 0; t < P.nterms; ++t) {
            if (!__any_sync(0xffffffffu, keep)) break;
            keep = keep && term_pass1<-1, false>(P.terms[t], stage, rr, meta.null_mask, P, meta.page);
          }
        }
        uint32_t m = __ballot_sync(0xffffffffu, keep);
        n_filt += __popc(m);
        if constexpr (!kNoNull) {
          // NULL keys never match (an inner join drops the row) and are DefinitelyAbsent for a filter (shared.rs:367-374)
          if (P.njoins | P.bloom_dense) {
            if (P.bloom_dense) {
              const uint32_t rej = __popc(__ballot_sync(0xffffffffu, keep && !kvalid));
              if (lane == 0) ctl->n_bloom_rej += rej;
            }
            keep = keep && kvalid;
            m = __ballot_sync(0xffffffffu, keep);
          }
        }
        if (m) {
          if (keep) q1[q1n + __popc(m & lt)] = make_uint4(meta.page, meta.r0 + rr, uint32_t(uint64_t(key)), uint32_t(uint64_t(key) >> 32));
          q1n += __popc(m);
          __syncwarp();
          if (q1n >= 32u) {
            run_b(32u);
            if (q2n >= 32u) run_c(32u);
          }
        }
      }
      __syncwarp();   // every lane is done reading the stage: refill it with the tile D ahead
      issue_next();
    }
    // drain: the rest of q1 goes through stage B, the last pending batch is resolved, the rest of q2 through stage C
    while (q1n | uint32_t(pend)) {
      run_b(min(q1n, 32u));
      if (q2n >= 32u) run_c(32u);
    }
    while (q2n) run_c(min(q2n, 32u));
    __syncwarp();
    const uint32_t n_in = ctl->n_in;
    if (P.nbloom == P.bloom_dense) n_bloom = n_in;   // no probe ran in stage A
    n_bloom -= ctl->n_bloom_rej;                     // rows_bloom = rows the runtime filters did not reject
    n_filt -= ctl->n_bloom_rej;                      // rows_filtered = rows past the filters AND the predicate
  }

  // counters (RuntimeFilter*/Worker* style metrics).  n_in .. n_out are warp-uniform in the consumer
  // warps (ballot popcounts): one lane adds them; n_bad and n_bloom_ins are per lane.
  {
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(P.counters);
    uint32_t bad = n_bad;
#pragma unroll
    for (int o = 16; o; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
    if (lane == 0) {
      bad += ctl->n_bad;
      if (ctl->n_in) atomicAdd(dst + 0, (unsigned long long)ctl->n_in);
      if (n_bloom) atomicAdd(dst + 1, (unsigned long long)n_bloom);
      if (n_filt) atomicAdd(dst + 2, (unsigned long long)n_filt);
      if (ctl->n_out) atomicAdd(dst + 3, (unsigned long long)ctl->n_out);
      if (bad) atomicAdd(dst + 5, (unsigned long long)bad);
      if (ctl->n_groups) atomicAdd(P.table.used, ctl->n_groups);   // groups this warp created
    }
  }
}

// ---- stage C as a kernel of its own (split execution: aggregate sinks behind ONE join) --------------
// 32 warps per SM.  Two phases per warp, both on dense lanes:
//   match   32 entries at a time: walk the probe chain (tag window from L2, slot from HBM, key compare) and push
//           every (row, matched slot) pair -- duplicates on the build side multiply -- into the warp's queue.  Two
//           thirds of the lineitem side's entries are false hits of the 8-bit tag and end here, ~100 instructions in;
//   sink    whenever 32 matches are queued: assemble the group key from page and payload columns, elect one leader per
//           group (match.any), look the group up, evaluate the aggregate arguments and add them.  This is the
//           expensive half (~900 instructions per batch) and it now always runs with full lanes.
// Nothing else waits for these chains: the scan kernel (stages A and B) has finished its tiles by then.
constexpr uint32_t kEntryQueue = 64;
template <uint32_t ACC>
__global__ void __launch_bounds__(256, 4) entries_pipeline_kernel(const __grid_constant__ DevPlan P) {
  __shared__ uint4 s_qe[8][kEntryQueue];             // matched rows {page, row, key lo, key hi}
  __shared__ const uint32_t* s_qs[8][kEntryQueue];   // ... and the slot each of them matched
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint4* qe = s_qe[warp];
  const uint32_t** qs = s_qs[warp];
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t qn = 0, n_out = 0, n_bad = 0, n_groups = 0;
  const DevJoin& j = P.joins[0];
  auto sink = [&](uint32_t n) {
    qn -= n;
    const bool act = lane < n;
    uint4 e = make_uint4(0, 0, 0, 0);
    GRow g;
    g.pay0 = g.pay1 = g.rec = nullptr;
    g.occ0 = g.occ1 = g.occr = 0;
    if (act) {
      e = qe[qn + lane];
      g.pay0 = qs[qn + lane];
      g.occ0 = __ldg(g.pay0 + 2);
    }
    __syncwarp();
    g.page = P.pages + uint64_t(e.x) * P.page_stride;
    g.r = e.y;
    if (P.single_class && !P.used_null_mask) {
      g.lc = &P.class0;
      g.nulls = 0;
    } else {
      PageDesc d{};
      if (act) d = P.descs[e.x];
      g.lc = P.single_class ? &P.class0 : P.classes + d.layout_class;
      g.nulls = d.null_mask & P.used_null_mask;
    }
    n_out += n;
    if (P.nkeys) sink_agg_grouped<ACC>(P, g, act, lane, n_bad, n_groups);
    else sink_agg_single<ACC>(P, g, act, lane);
    __syncwarp();
  };
  const unsigned long long appended = *P.entries_count;
  const uint64_t n = appended < P.entries_cap ? appended : P.entries_cap;
  const uint64_t nwarps = uint64_t(gridDim.x) * (blockDim.x >> 5), w0 = uint64_t(blockIdx.x) * (blockDim.x >> 5) + warp;
  // two entries per lane and step: the two chains (tag window from L2, then the slot from HBM) are independent, so
  // their round trips overlap -- the match phase is nothing but latency
  for (uint64_t base = w0 * 64u; base < n; base += nwarps * 64u) {
    const uint64_t ia = base + lane, ib = ia + 32u;
    const bool acta = ia < n, actb = ib < n;
    uint4 ea = make_uint4(0, 0, 0, 0), eb = make_uint4(0, 0, 0, 0);
    if (acta) ea = __ldg(P.entries + ia);
    if (actb) eb = __ldg(P.entries + ib);
    JoinIter ita, itb;
    ita.init(j, int64_t((uint64_t(ea.w) << 32) | ea.z), acta);   // (NULL keys never reach the entries)
    itb.init(j, int64_t((uint64_t(eb.w) << 32) | eb.z), actb);
    for (;;) {
      // both chains advance to their next candidate slot, both slots are requested, then both keys are compared
      const uint4* ca = ita.next_candidate(j);
      const uint4* cb = itb.next_candidate(j);
      if (!__any_sync(0xffffffffu, (ca != nullptr) | (cb != nullptr))) break;
      uint4 sa = make_uint4(0, 0, 0, 0), sb = make_uint4(0, 0, 0, 0);
      if (ca) sa = __ldg(ca);
      if (cb) sb = __ldg(cb);
      const bool fa = ca && sa.x == ita.klo && sa.y == ita.khi, fb = cb && sb.x == itb.klo && sb.y == itb.khi;
      const uint32_t ma = __ballot_sync(0xffffffffu, fa), mb = __ballot_sync(0xffffffffu, fb);
      if (ma) {
        if (fa) {
          const uint32_t at = qn + __popc(ma & lt);
          qe[at] = ea;
          qs[at] = reinterpret_cast<const uint32_t*>(ca);
        }
        qn += __popc(ma);
        __syncwarp();
        if (qn >= 32u) sink(32u);
      }
      if (mb) {
        if (fb) {
          const uint32_t at = qn + __popc(mb & lt);
          qe[at] = eb;
          qs[at] = reinterpret_cast<const uint32_t*>(cb);
        }
        qn += __popc(mb);
        __syncwarp();
        if (qn >= 32u) sink(32u);
      }
    }
  }
  if (qn) sink(qn);
  unsigned long long* dst = reinterpret_cast<unsigned long long*>(P.counters);
  if (lane == 0 && n_out) atomicAdd(dst + 3, (unsigned long long)n_out);   // (warp uniform)
  uint32_t vals[2] = {n_groups, n_bad};
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    uint32_t v = vals[q];
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && v) {
      if (q == 0) atomicAdd(P.table.used, v);
      else atomicAdd(dst + 5, (unsigned long long)v);
    }
  }
}

// ---- row-set scans ------------------------------------------------------------------------------
// The input is a dense array of rows in build-row format {key lo, key hi, occupancy / NULL flags, payload...}
// -- what a build sink emits and what a hash-partitioned exchange delivers -- instead of pages: every warp takes
// 32 rows at a time straight into stage C (probe chains, sink).  Column 0 of the scan is the key, column i + 1
// payload i (DevRef.src == kSrcRecord).
template <uint32_t ACC>
__global__ void __launch_bounds__(256) rows_pipeline_kernel(const __grid_constant__ DevPlan P) {
  const uint32_t lane = threadIdx.x & 31;
  uint32_t n_out = 0, n_bad = 0, n_groups = 0, n_in = 0;
  const uint64_t nwarps = uint64_t(gridDim.x) * (blockDim.x >> 5), w0 = uint64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  for (uint64_t base = w0 * 32u; base < P.row_count; base += nwarps * 32u) {
    const uint64_t i = base + lane;
    const bool act = i < P.row_count;
    uint4 e = make_uint4(uint32_t(i), uint32_t(i >> 32), 0u, 0u);
    if (act) {
      const uint4 head = __ldg(P.row_src + i * P.row_u4);
      e.z = head.x;
      e.w = head.y;
    }
    n_in += __popc(__ballot_sync(0xffffffffu, act));
    stage_c<ACC, true>(P, e, act, lane, n_out, n_bad, n_groups);
  }
  unsigned long long* dst = reinterpret_cast<unsigned long long*>(P.counters);
  if (lane == 0) {
    if (n_in) { atomicAdd(dst + 0, (unsigned long long)n_in); atomicAdd(dst + 1, (unsigned long long)n_in); atomicAdd(dst + 2, (unsigned long long)n_in); }
    if (n_out) atomicAdd(dst + 3, (unsigned long long)n_out);
  }
  uint32_t vals[2] = {n_groups, n_bad};
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    uint32_t v = vals[q];
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && v) {
      if (q == 0) atomicAdd(P.table.used, v);
      else atomicAdd(dst + 5, (unsigned long long)v);
    }
  }
}

}  // namespace pgf
