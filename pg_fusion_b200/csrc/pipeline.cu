// Host side of the fused pipelines: lowering of the C-ABI plan (pgf_pipeline) into the
// device plan, eligibility rules, launch, and extraction of AggregateExec results.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "context.hpp"
#include "layout.hpp"
#include "pipeline_kernel.cuh"
#include "pipeline_shapes.hpp"
#include "probe_kernel.cuh"

namespace pgf {

// implemented in pipeline_inst_*.cu (explicit instantiations split for parallel builds)
cudaError_t launch_pipeline(uint32_t sink, uint32_t acc, bool grouped, uint32_t nj, uint32_t maxe,
                            const DevPlan& plan, uint32_t grid, size_t smem, cudaStream_t stream);
// implemented in pipeline_inst_probe.cu: the compaction pipeline (joins, build sinks)
cudaError_t launch_probe(uint32_t acc, int t0, const DevPlan& plan, uint32_t grid, size_t smem, cudaStream_t stream);
cudaError_t launch_probe_split(uint32_t acc, int t0, const DevPlan& plan, uint32_t grid, size_t smem, uint32_t cgrid, cudaStream_t stream);
cudaError_t launch_rows(uint32_t acc, const DevPlan& plan, uint32_t grid, cudaStream_t stream);

namespace {

// PGF_TRACE=1: synchronise at phase boundaries of pgf_pipeline_run and print the elapsed host time of
// each phase to stderr (development aid; off by default, adds synchronisations when on).
struct PhaseTrace {
  bool on = false;
  cudaStream_t stream;
  std::chrono::steady_clock::time_point t;
  explicit PhaseTrace(cudaStream_t s) : stream(s) {
    const char* e = std::getenv("PGF_TRACE");
    on = e && *e && *e != '0';
    if (const char* only = std::getenv("PGF_TRACE_RANK")) {   // one rank of a torchrun job
      const char* r = std::getenv("RANK");
      on = on && r && std::strcmp(r, only) == 0;
    }
    if (on) { cudaStreamSynchronize(stream); t = std::chrono::steady_clock::now(); }
  }
  void mark(const char* what) {
    if (!on) return;
    cudaStreamSynchronize(stream);
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[pgf trace] %-28s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
    t = now;
  }
};

// ---- device helpers for the group table ------------------------------------------------
// Partial-state entry: [kKeyWords key][1 null mask][nexprs * acc_words acc][nexprs + 1 counts]
__host__ __device__ inline uint32_t entry_words(uint32_t nexprs, uint32_t acc_words) {
  return kKeyWords + 1 + nexprs * acc_words + nexprs + 1;
}

// Compacts the occupied slots into out = [count][entries...]; out[0] must be zero on entry.
// An aggregate without GROUP BY always has its single output row (slot 0).  Every block owns a
// contiguous segment of the table: it counts its occupied slots, reserves its output range with ONE
// atomic (a per-slot atomic on the single counter serialises: 26 ms for 1.1 M groups) and writes
// its entries at block-local prefix positions.
constexpr uint32_t kExtractThreads = 256;
__device__ __forceinline__ void extract_entry(const GroupTable& t, uint32_t nexprs, bool grouped, uint64_t i, uint32_t st, uint64_t* e) {
  for (uint32_t w = 0; w < kKeyWords; ++w) e[w] = t.keys[i * kKeyWords + w];
  e[kKeyWords] = grouped ? st >> 8 : 0u;
  for (uint32_t w = 0; w < nexprs * t.acc_words; ++w) e[kKeyWords + 1 + w] = t.acc[i * nexprs * t.acc_words + w];
  for (uint32_t w = 0; w <= nexprs; ++w) e[kKeyWords + 1 + nexprs * t.acc_words + w] = t.cnt[i * (nexprs + 1) + w];
}
__global__ void __launch_bounds__(kExtractThreads) table_extract_kernel(GroupTable t, uint32_t nexprs, bool grouped, uint64_t* out,
                                                                       uint64_t max_entries) {
  // Every warp owns a contiguous run of slots: it counts its occupied slots (four state words per lane and load),
  // the block turns the eight counts into offsets behind ONE reservation, and the warp then walks its run again
  // placing entries by a shuffle prefix over the lanes -- no block-wide barrier per chunk (the first version had two
  // per 256 slots: 0.36 ms for the 32 M-slot table of Q3 at SF100).
  __shared__ unsigned long long s_base;
  __shared__ uint32_t s_warp[kExtractThreads / 32];
  constexpr uint32_t kWarps = kExtractThreads / 32;
  const uint32_t ew = entry_words(nexprs, t.acc_words);
  if (!grouped) {  // an aggregate without GROUP BY always has its single output row (slot 0)
    if (blockIdx.x == 0 && threadIdx.x == 0 && max_entries) {
      out[0] = 1;
      extract_entry(t, nexprs, false, 0, 0u, out + 1);
    }
    return;
  }
  const uint64_t nslots = uint64_t(t.mask) + 1;   // a power of two >= 1024
  const uint64_t per_warp = ((nslots + uint64_t(gridDim.x) * kWarps - 1) / (uint64_t(gridDim.x) * kWarps) + 127) / 128 * 128;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t w0 = (uint64_t(blockIdx.x) * kWarps + warp) * per_warp, w1 = w0 + per_warp < nslots ? w0 + per_warp : nslots;
  auto states = [&](uint64_t c0) {  // the four state words of this lane in the 128-slot chunk at c0 (0 beyond the run)
    const uint64_t i = c0 + lane * 4u;
    return i < w1 ? *reinterpret_cast<const uint4*>(t.state + i) : make_uint4(0u, 0u, 0u, 0u);
  };
  auto occ = [](uint32_t s) { return uint32_t((s & 3u) == 2u); };
  uint32_t mine = 0;
  for (uint64_t c0 = w0; c0 < w1; c0 += 128) {
    const uint4 s = states(c0);
    mine += occ(s.x) + occ(s.y) + occ(s.z) + occ(s.w);
  }
  mine = __reduce_add_sync(0xffffffffu, mine);
  if (lane == 0) s_warp[warp] = mine;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t total = 0;
    for (uint32_t w = 0; w < kWarps; ++w) total += s_warp[w];
    s_base = total ? atomicAdd(reinterpret_cast<unsigned long long*>(out), (unsigned long long)total) : 0ull;
  }
  __syncthreads();
  unsigned long long pos = s_base;
  for (uint32_t w = 0; w < warp; ++w) pos += s_warp[w];
  if (!mine) return;
  for (uint64_t c0 = w0; c0 < w1; c0 += 128) {
    const uint4 s = states(c0);
    const uint32_t st[4] = {s.x, s.y, s.z, s.w};
    const uint32_t n = occ(s.x) + occ(s.y) + occ(s.z) + occ(s.w);
    uint32_t incl = n;   // inclusive prefix of n over the lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (int(lane) >= o) incl += v;
    }
    unsigned long long at = pos + (incl - n);
#pragma unroll
    for (uint32_t q = 0; q < 4; ++q) {
      if (occ(st[q])) {
        if (at < max_entries) extract_entry(t, nexprs, true, c0 + lane * 4u + q, st[q], out + 1 + at * ew);
        ++at;
      }
    }
    pos += __shfl_sync(0xffffffffu, incl, 31);
  }
}

// ---- broadcast-join exchange: a built table travels between GPUs as its occupied
// slots.  Slots are self-contained records {key lo, key hi, occupancy/NULL flags, payload...}.
// (block-aggregated like table_extract_kernel: one atomic per block reserves the output range)
__global__ void __launch_bounds__(kExtractThreads) join_export_kernel(const uint4* slots, const uint8_t* tags, uint32_t capacity, uint32_t slot_u4,
                                                                     uint4* out, unsigned long long max_rows, unsigned long long* count) {
  __shared__ unsigned long long s_base;
  __shared__ uint32_t s_warp[kExtractThreads / 32];
  const uint64_t seg = ((uint64_t(capacity) + gridDim.x - 1) / gridDim.x + kExtractThreads - 1) / kExtractThreads * kExtractThreads;
  const uint64_t b0 = uint64_t(blockIdx.x) * seg, b1 = b0 + seg < capacity ? b0 + seg : capacity;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  auto occupied = [&](uint64_t i) { return i < b1 && tags[i] != 0; };   // the slot array is never cleared: the tag says whether a slot is live
  uint32_t mine = 0;
  for (uint64_t i = b0 + threadIdx.x; i < b1; i += kExtractThreads) mine += occupied(i);
  mine = __reduce_add_sync(0xffffffffu, mine);
  if (lane == 0) s_warp[warp] = mine;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t total = 0;
    for (uint32_t w = 0; w < kExtractThreads / 32; ++w) total += s_warp[w];
    s_base = total ? atomicAdd(count, (unsigned long long)total) : 0ull;
  }
  __syncthreads();
  unsigned long long pos_base = s_base;
  for (uint64_t c0 = b0; c0 < b1; c0 += kExtractThreads) {
    const uint64_t i = c0 + threadIdx.x;
    const bool occ = occupied(i);
    const uint32_t mask = __ballot_sync(0xffffffffu, occ);
    __syncthreads();
    if (lane == 0) s_warp[warp] = __popc(mask);
    __syncthreads();
    uint32_t before = 0, chunk_total = 0;
    for (uint32_t w = 0; w < kExtractThreads / 32; ++w) {
      if (w < warp) before += s_warp[w];
      chunk_total += s_warp[w];
    }
    if (occ) {
      const unsigned long long pos = pos_base + before + __popc(mask & ((1u << lane) - 1u));
      if (pos < max_rows) {
        out[pos * slot_u4] = slots[i * slot_u4];
        if (slot_u4 == 2) out[pos * 2 + 1] = slots[i * 2 + 1];
      }
    }
    pos_base += chunk_total;
  }
}

// Builds the table from dense build rows (the output of a build-sink pipeline, or the fragments of a broadcast /
// partitioned exchange).  A row claims the first empty tag byte of its home bucket (or of the buckets that follow:
// linear probing) with a 64-bit CAS on the bucket's tag word, then writes its slot.  Only the tag directory has to
// start out zeroed and only it sees atomics -- 1 byte per slot, L2 resident -- while the slot array is written
// exactly once, a whole 16 / 32-byte record per row, and never cleared.  capacity >= 2 x rows, so an empty tag
// always exists.
__global__ void join_build_kernel(uint4* slots, uint8_t* tags, uint32_t mask, uint32_t shift, uint32_t slot_u4, const uint4* rows, uint64_t nrows) {
  for (uint64_t r = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; r < nrows; r += uint64_t(gridDim.x) * blockDim.x) {
    const uint4 s0 = rows[r * slot_u4];
    const uint64_t hk = join_hash(int64_t((uint64_t(s0.y) << 32) | s0.x));
    const unsigned long long tag = join_tag8(hk, shift);
    uint32_t base = join_home(hk, shift);
    for (;;) {
      unsigned long long* word = reinterpret_cast<unsigned long long*>(tags + base);
      unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(word);
      uint32_t pos = kJoinBucket;
      for (;;) {
        const uint32_t z0 = ~(((uint32_t(cur) & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | uint32_t(cur) | 0x7F7F7F7Fu);          // 0x80 per empty tag
        const uint32_t z1 = ~(((uint32_t(cur >> 32) & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | uint32_t(cur >> 32) | 0x7F7F7F7Fu);
        if (!(z0 | z1)) break;                                                      // bucket full: go on with the next one
        pos = z0 ? (uint32_t(__ffs(int(z0))) - 1u) >> 3 : 4u + ((uint32_t(__ffs(int(z1))) - 1u) >> 3);
        const unsigned long long seen = atomicCAS(word, cur, cur | (tag << (8u * pos)));
        if (seen == cur) break;                                                     // the tag byte is ours
        cur = seen;
        pos = kJoinBucket;
      }
      if (pos < kJoinBucket) {
        const uint64_t i = uint64_t(base) + pos;
        slots[i * slot_u4] = s0;
        if (slot_u4 == 2) slots[i * 2 + 1] = rows[r * 2 + 1];
        break;
      }
      base = (base + kJoinBucket) & mask;
    }
  }
}

// RuntimeFilterBuildExec (worker_runtime/src/runtime_filter_plan.rs:227-274,345-363): every build row's key goes into
// the filter.  Run over the dense build rows after the fused kernel: full lanes, coalesced key reads, and the atomics
// on the bit array are not tangled with the probe chains of stage C.  (NULL keys never reach the rows.)
__global__ void bloom_insert_rows_kernel(DevBloom b, const uint4* rows, uint64_t nrows, uint32_t row_u4) {
  for (uint64_t r = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; r < nrows; r += uint64_t(gridDim.x) * blockDim.x) {
    const uint4 s0 = rows[r * row_u4];
    bloom_insert(b, (uint64_t(s0.y) << 32) | s0.x);
  }
}

// ---- hash-partitioned exchange: rows grouped by owner rank = join_partition(key) --------
// pass 1 counts the rows of every destination, pass 2 scatters them behind per-destination cursors (block-local
// counts first, one atomic per block and destination).
constexpr uint32_t kMaxRanks = 16;
__global__ void partition_count_kernel(const uint4* rows, uint64_t nrows, uint32_t row_u4, uint32_t world, unsigned long long* counts) {
  __shared__ uint32_t s_cnt[kMaxRanks];
  if (threadIdx.x < kMaxRanks) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  for (uint64_t r = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; r < nrows; r += uint64_t(gridDim.x) * blockDim.x) {
    const uint4 s0 = rows[r * row_u4];
    atomicAdd(&s_cnt[join_partition(int64_t((uint64_t(s0.y) << 32) | s0.x), world)], 1u);
  }
  __syncthreads();
  if (threadIdx.x < world && s_cnt[threadIdx.x]) atomicAdd(counts + threadIdx.x, (unsigned long long)s_cnt[threadIdx.x]);
}
__global__ void partition_scatter_kernel(const uint4* rows, uint64_t nrows, uint32_t row_u4, uint32_t world, unsigned long long* cursors,
                                         uint4* out) {
  __shared__ uint32_t s_cnt[kMaxRanks];
  __shared__ unsigned long long s_base[kMaxRanks];
  const uint64_t per_block = (nrows + gridDim.x - 1) / gridDim.x;
  const uint64_t b0 = blockIdx.x * per_block, b1 = b0 + per_block < nrows ? b0 + per_block : nrows;
  if (threadIdx.x < kMaxRanks) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  for (uint64_t r = b0 + threadIdx.x; r < b1; r += blockDim.x) {
    const uint4 s0 = rows[r * row_u4];
    atomicAdd(&s_cnt[join_partition(int64_t((uint64_t(s0.y) << 32) | s0.x), world)], 1u);
  }
  __syncthreads();
  if (threadIdx.x < world) {
    s_base[threadIdx.x] = s_cnt[threadIdx.x] ? atomicAdd(cursors + threadIdx.x, (unsigned long long)s_cnt[threadIdx.x]) : 0ull;
    s_cnt[threadIdx.x] = 0;
  }
  __syncthreads();
  for (uint64_t r = b0 + threadIdx.x; r < b1; r += blockDim.x) {
    const uint4 s0 = rows[r * row_u4];
    const uint32_t d = join_partition(int64_t((uint64_t(s0.y) << 32) | s0.x), world);
    const unsigned long long pos = s_base[d] + atomicAdd(&s_cnt[d], 1u);
    out[pos * row_u4] = s0;
    if (row_u4 == 2) out[pos * 2 + 1] = rows[r * 2 + 1];
  }
}

// ---- SortExec / TopK above the aggregate -------------------------------------------------
// ORDER BY over the extracted group entries.  With a small LIMIT the k best entries are
// selected on the device (k rounds of a parallel arg-best reduction under the ORDER BY
// comparator) and only those leave the GPU; the host puts the k rows in their final order.
enum : uint32_t { SK_KEY_INT = 0, SK_KEY_VIEW = 1, SK_KEY_DEC = 2, SK_F64_SUM = 3, SK_F64_AVG = 4, SK_I64_SUM = 5,
                  SK_I128_SUM = 6, SK_I128_AVG = 7, SK_COUNT = 8 };
struct DevSortKey {
  uint32_t kind, word, cnt_word, null_bit, desc, nulls_first;
};
struct DevSort {
  uint32_t n, ew;
  DevSortKey k[PGF_MAX_SORT];
};

__device__ __forceinline__ int cmp_u64(uint64_t a, uint64_t b) { return a < b ? -1 : (a > b ? 1 : 0); }
__device__ __forceinline__ int cmp_i64(int64_t a, int64_t b) { return a < b ? -1 : (a > b ? 1 : 0); }
__device__ __forceinline__ int cmp_f64_total(double a, double b) {  // IEEE totalOrder, as arrow's sort
  int64_t x = __double_as_longlong(a), y = __double_as_longlong(b);
  x ^= int64_t(uint64_t(x >> 63) >> 1);
  y ^= int64_t(uint64_t(y >> 63) >> 1);
  return cmp_i64(x, y);
}
__device__ __forceinline__ int cmp_i128(const uint64_t* a, const uint64_t* b) {
  const int c = cmp_i64(int64_t(a[1]), int64_t(b[1]));
  return c ? c : cmp_u64(a[0], b[0]);
}
__device__ __forceinline__ int cmp_view(const uint64_t* a, const uint64_t* b) {  // bytes, then length
  auto be = [](uint32_t v) { return __byte_perm(v, 0, 0x0123); };
  const uint32_t pa[3] = {be(uint32_t(a[0] >> 32)), be(uint32_t(a[1])), be(uint32_t(a[1] >> 32))};
  const uint32_t pb[3] = {be(uint32_t(b[0] >> 32)), be(uint32_t(b[1])), be(uint32_t(b[1] >> 32))};
  for (int i = 0; i < 3; ++i)
    if (pa[i] != pb[i]) return pa[i] < pb[i] ? -1 : 1;
  return cmp_u64(uint32_t(a[0]), uint32_t(b[0]));
}

// < 0: entry a sorts before entry b.  Ties are broken by the raw key words so the order is total.
__device__ int sort_compare(const DevSort& S, const uint64_t* a, const uint64_t* b) {
  for (uint32_t i = 0; i < S.n; ++i) {
    const DevSortKey& k = S.k[i];
    bool na, nb;
    if (k.kind <= SK_KEY_DEC) {
      na = (a[kKeyWords] >> k.null_bit) & 1;
      nb = (b[kKeyWords] >> k.null_bit) & 1;
    } else if (k.kind == SK_COUNT) {
      na = nb = false;
    } else {
      na = a[k.cnt_word] == 0;
      nb = b[k.cnt_word] == 0;
    }
    if (na || nb) {
      if (na && nb) continue;
      return (na == bool(k.nulls_first)) ? -1 : 1;
    }
    int c;
    switch (k.kind) {
      case SK_KEY_INT: c = cmp_i64(int64_t(a[k.word]), int64_t(b[k.word])); break;
      case SK_KEY_VIEW: c = cmp_view(a + k.word, b + k.word); break;
      case SK_KEY_DEC: c = cmp_i128(a + k.word, b + k.word); break;
      case SK_F64_SUM: c = cmp_f64_total(__longlong_as_double(a[k.word]), __longlong_as_double(b[k.word])); break;
      case SK_F64_AVG:
        c = cmp_f64_total(__longlong_as_double(a[k.word]) / double(a[k.cnt_word]), __longlong_as_double(b[k.word]) / double(b[k.cnt_word]));
        break;
      case SK_I64_SUM: c = cmp_i64(int64_t(a[k.word]), int64_t(b[k.word])); break;
      case SK_I128_SUM: c = cmp_i128(a + k.word, b + k.word); break;
      case SK_I128_AVG: {
        const __int128 x = (__int128)(((unsigned __int128)a[k.word + 1] << 64) | a[k.word]) * 10000 / (__int128)a[k.cnt_word];
        const __int128 y = (__int128)(((unsigned __int128)b[k.word + 1] << 64) | b[k.word]) * 10000 / (__int128)b[k.cnt_word];
        c = x < y ? -1 : (x > y ? 1 : 0);
        break;
      }
      default: c = cmp_u64(a[k.cnt_word], b[k.cnt_word]); break;  // COUNT
    }
    if (c) return k.desc ? -c : c;
  }
  for (uint32_t w = 0; w <= kKeyWords; ++w)
    if (a[w] != b[w]) return a[w] < b[w] ? -1 : 1;
  return 0;
}

constexpr uint32_t kTopkThreads = 256;
constexpr uint32_t kTopkItems = 8;                              // entries per thread
constexpr uint32_t kTopkSeg = kTopkThreads * kTopkItems;        // entries per block and level
constexpr uint32_t kNoEntry = 0xFFFFFFFFu;

// Order-preserving 64-bit summary of an entry under the FIRST ORDER BY term (smaller sorts first):
//   summary(a) < summary(b)  implies  a sorts before b;
// equal summaries say nothing and are decided by sort_compare().  One or two loads and a handful of integer
// instructions per entry, against the generic comparator's loop over terms, kinds and tie-break words.
__device__ uint64_t sort_summary(const DevSort& S, const uint64_t* e) {
  const DevSortKey& k = S.k[0];
  bool null;
  if (k.kind <= SK_KEY_DEC) null = (e[kKeyWords] >> k.null_bit) & 1;
  else if (k.kind == SK_COUNT) null = false;
  else null = e[k.cnt_word] == 0;
  if (null) return k.nulls_first ? 0ull : ~0ull;
  constexpr uint64_t kSign = 0x8000000000000000ull;
  auto sat64 = [&](uint64_t lo, uint64_t hi) -> uint64_t {   // i128 -> i64, saturating: monotone, exact for the values that fit
    if (hi == uint64_t(int64_t(lo) >> 63)) return lo ^ kSign;
    return int64_t(hi) < 0 ? 0ull : ~0ull;
  };
  auto f64key = [&](double d) -> uint64_t {                  // IEEE totalOrder, as an unsigned key
    int64_t x = __double_as_longlong(d);
    x ^= int64_t(uint64_t(x >> 63) >> 1);
    return uint64_t(x) ^ kSign;
  };
  uint64_t v;
  switch (k.kind) {
    case SK_KEY_INT: case SK_I64_SUM: v = e[k.word] ^ kSign; break;
    case SK_KEY_VIEW:   // the first eight bytes of the string, big-endian
      v = (uint64_t(__byte_perm(uint32_t(e[k.word] >> 32), 0, 0x0123)) << 32) | __byte_perm(uint32_t(e[k.word + 1]), 0, 0x0123);
      break;
    case SK_KEY_DEC: case SK_I128_SUM: v = sat64(e[k.word], e[k.word + 1]); break;
    case SK_F64_SUM: v = f64key(__longlong_as_double(e[k.word])); break;
    case SK_F64_AVG: v = f64key(__longlong_as_double(e[k.word]) / double(e[k.cnt_word])); break;
    case SK_I128_AVG: {
      const __int128 q = (__int128)(((unsigned __int128)e[k.word + 1] << 64) | e[k.word]) * 10000 / (__int128)e[k.cnt_word];
      v = sat64(uint64_t(q), uint64_t(q >> 64));
      break;
    }
    default: v = e[k.cnt_word]; break;  // COUNT
  }
  if (k.desc) v = ~v;
  return v < 1ull ? 1ull : (v > ~0ull - 1ull ? ~0ull - 1ull : v);   // 0 and ~0 belong to the NULLs
}

struct TopkPick {
  uint64_t p;
  uint32_t i;
};
__device__ __forceinline__ TopkPick topk_better(const DevSort& S, const uint64_t* entries, TopkPick x, TopkPick y) {
  if (x.i == kNoEntry) return y;
  if (y.i == kNoEntry) return x;
  if (x.p != y.p) return x.p < y.p ? x : y;
  return sort_compare(S, entries + uint64_t(x.i) * S.ew, entries + uint64_t(y.i) * S.ew) <= 0 ? x : y;
}

// Top-k selection (k <= PGF_TOPK_DEVICE_MAX), one level: every block takes kTopkSeg entries (the table's entries at
// level 0, the candidates of the level below afterwards), keeps their summaries in registers and runs k rounds of a
// block-wide arg-best (thread-local scan of 8 registers, shuffle butterfly, 8 warp winners).  A level shrinks its
// input by kTopkSeg / k >= 32, so 1.1 M groups take three launches of a few microseconds each (the first version --
// k rounds of the generic comparator over global memory, then ONE block over every block's candidates -- took
// 0.5 ms at SF10, longer than the lineitem pipeline itself).  out_rows != nullptr: the final level; the winners'
// entries are written in order, out_rows[0] = how many.
__global__ void __launch_bounds__(kTopkThreads) topk_select_kernel(DevSort S, const uint64_t* entries, const uint32_t* in_idx, uint32_t n, uint32_t k,
                                                                    uint32_t* out_idx, uint64_t* out_rows) {
  __shared__ uint64_t s_p[kTopkThreads / 32];
  __shared__ uint32_t s_i[kTopkThreads / 32];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t b0 = uint64_t(blockIdx.x) * kTopkSeg;
  TopkPick mine[kTopkItems];
#pragma unroll
  for (uint32_t j = 0; j < kTopkItems; ++j) {
    const uint64_t at = b0 + j * kTopkThreads + threadIdx.x;
    uint32_t i = kNoEntry;
    if (at < n) i = in_idx ? in_idx[at] : uint32_t(at);
    mine[j].i = i;
    mine[j].p = i != kNoEntry ? sort_summary(S, entries + uint64_t(i) * S.ew) : ~0ull;
  }
  uint32_t found = 0;
  for (uint32_t r = 0; r < k; ++r) {
    TopkPick best{~0ull, kNoEntry};
#pragma unroll
    for (uint32_t j = 0; j < kTopkItems; ++j) best = topk_better(S, entries, best, mine[j]);
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const TopkPick other{__shfl_xor_sync(0xffffffffu, best.p, o), __shfl_xor_sync(0xffffffffu, best.i, o)};
      best = topk_better(S, entries, best, other);   // (a total order: both lanes of a pair keep the same entry)
    }
    if (lane == 0) { s_p[warp] = best.p; s_i[warp] = best.i; }
    __syncthreads();
    best = TopkPick{s_p[0], s_i[0]};
#pragma unroll
    for (uint32_t w = 1; w < kTopkThreads / 32; ++w) best = topk_better(S, entries, best, TopkPick{s_p[w], s_i[w]});
    __syncthreads();
    if (threadIdx.x == 0) out_idx[uint64_t(blockIdx.x) * k + r] = best.i;
    if (best.i == kNoEntry) {   // fewer than k entries in this segment
      for (uint32_t q = r + 1 + threadIdx.x; q < k; q += kTopkThreads) out_idx[uint64_t(blockIdx.x) * k + q] = kNoEntry;
      break;
    }
    if (out_rows)
      for (uint32_t w = threadIdx.x; w < S.ew; w += kTopkThreads) out_rows[1 + uint64_t(r) * S.ew + w] = entries[uint64_t(best.i) * S.ew + w];
    found = r + 1;
#pragma unroll
    for (uint32_t j = 0; j < kTopkItems; ++j)
      if (mine[j].i == best.i) mine[j].i = kNoEntry;   // taken
  }
  if (out_rows && threadIdx.x == 0) out_rows[0] = found;
}

// Final merge of partial states (AggregateExec FinalPartitioned): one launch per state, in
// rank order, so Float64 sums are added in a fixed order on every rank.
template <uint32_t ACC>
__global__ void table_merge_kernel(GroupTable t, uint32_t nexprs, uint32_t nkeywords, const uint64_t* state, uint64_t max_entries) {
  using Ops = AccOps<ACC>;
  uint64_t n = state[0];
  if (n > max_entries) {  // the producing rank had more groups than its state buffer holds
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicExch(t.overflow, 2u);
    n = max_entries;
  }
  const uint32_t ew = entry_words(nexprs, t.acc_words);
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x) {
    const uint64_t* e = state + 1 + i * ew;
    const int64_t slot = nkeywords ? group_slot(t, e, nkeywords, uint32_t(e[kKeyWords])) : 0;
    if (slot < 0) continue;
    for (uint32_t x = 0; x < nexprs; ++x) {
      typename Ops::T v;
      memcpy(&v, e + kKeyWords + 1 + x * t.acc_words, sizeof v);
      Ops::atomic_add(t.acc + (uint64_t(slot) * nexprs + x) * t.acc_words, v);
    }
    for (uint32_t x = 0; x <= nexprs; ++x)
      atomicAdd(reinterpret_cast<unsigned long long*>(t.cnt + uint64_t(slot) * (nexprs + 1) + x),
                (unsigned long long)e[kKeyWords + 1 + nexprs * t.acc_words + x]);
  }
}

// Small states (the bounded multi-GPU step): one CTA merges every rank's state, rank after rank
// with a barrier in between, so the fixed summation order costs one launch instead of one per rank.
template <uint32_t ACC>
__global__ void table_merge_all_kernel(GroupTable t, uint32_t nexprs, uint32_t nkeywords, const uint8_t* states, uint64_t stride,
                                       uint32_t nstates, uint64_t max_entries) {
  using Ops = AccOps<ACC>;
  const uint32_t ew = entry_words(nexprs, t.acc_words);
  for (uint32_t s = 0; s < nstates; ++s) {
    const uint64_t* state = reinterpret_cast<const uint64_t*>(states + s * stride);
    uint64_t n = state[0];
    if (n > max_entries) {
      if (threadIdx.x == 0) atomicExch(t.overflow, 2u);
      n = max_entries;
    }
    for (uint64_t i = threadIdx.x; i < n; i += blockDim.x) {
      const uint64_t* e = state + 1 + i * ew;
      const int64_t slot = nkeywords ? group_slot(t, e, nkeywords, uint32_t(e[kKeyWords])) : 0;
      if (slot < 0) continue;
      for (uint32_t x = 0; x < nexprs; ++x) {
        typename Ops::T v;
        memcpy(&v, e + kKeyWords + 1 + x * t.acc_words, sizeof v);
        Ops::atomic_add(t.acc + (uint64_t(slot) * nexprs + x) * t.acc_words, v);
      }
      for (uint32_t x = 0; x <= nexprs; ++x)
        atomicAdd(reinterpret_cast<unsigned long long*>(t.cnt + uint64_t(slot) * (nexprs + 1) + x),
                  (unsigned long long)e[kKeyWords + 1 + nexprs * t.acc_words + x]);
    }
    __threadfence();
    __syncthreads();
  }
}

// ---- lowering ----------------------------------------------------------------------
struct Lowered {
  DevPlan dev{};
  Scan* scan = nullptr;
  uint32_t acc_cls = CLS_F64;
  bool grouped = false;
  uint32_t nj = 0;
  uint32_t maxe = 2;
  size_t smem = 0;
  bool rowscan = false;     // scans a row set (rows_pipeline_kernel) instead of pages
  Scan row_scan;            // its stand-in scan: schema = [key, payload...], rows = rows of the set
  uint32_t bloom_dropped = 0;  // fused probes the lowering dropped: filter not Ready / other generation / redundant
  bool probe = false;       // runs the compaction pipeline (probe_kernel.cuh): joins and build sinks
  bool split = false;       // ... as two kernels: stages A + B, then stage C over the tag hits (plans with a join)
  uint32_t split_grid = 0;
  int t0 = -1;              // its predicate specialisation (LD_* of the single plain range term), -1 = generic
  int32_t key_types[4] = {0, 0, 0, 0};
  bool key_not_null[4] = {false, false, false, false};
  uint32_t expr_pos[kMaxExprs] = {0, 1, 2, 3, 4, 5, 6, 7};  // device position of the caller's expression e
  JoinTable build_table{};
  uint64_t table_capacity = 0;
};

struct DevAlloc {  // frees on scope exit unless released
  std::vector<void*> ptrs;
  ~DevAlloc() { for (void* p : ptrs) cudaFree(p); }
  template <class T>
  cudaError_t alloc(T** out, size_t count) {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, count * sizeof(T) ? count * sizeof(T) : 1);
    if (e == cudaSuccess) { ptrs.push_back(p); *out = static_cast<T*>(p); }
    return e;
  }
  void release(void* p) { ptrs.erase(std::remove(ptrs.begin(), ptrs.end(), p), ptrs.end()); }
};

bool is_int_type(int t) { return t == PGF_T_INT16 || t == PGF_T_INT32 || t == PGF_T_INT64; }
uint32_t type_u32_words(int t) {
  switch (t) {
    case PGF_T_INT16: case PGF_T_INT32: case PGF_T_FLOAT32: return 1;
    case PGF_T_INT64: case PGF_T_FLOAT64: return 2;
    case PGF_T_UTF8VIEW: case PGF_T_BINARYVIEW: case PGF_T_DECIMAL128: case PGF_T_UUID: return 4;
    default: return 0;
  }
}

const ShapeEntry* pick_shape(const Lowered& L);

class Lowering {
 public:
  Lowering(pgf_ctx* ctx, const pgf_pipeline* plan) : ctx_(ctx), plan_(plan) {}

  pgf_status run(Lowered* out) {
    L_ = out;
    if (plan_->scan_row_set) {
      // the pipeline reads a row set (the output of a build sink with PGF_BUILD_ROWS_ONLY or of an exchange):
      // column 0 is the key, column i + 1 payload i
      auto rt = ctx_->joins.find(plan_->scan_row_set);
      if (rt == ctx_->joins.end() || !rt->second.d_rows) return ctx_->fail(PGF_ERR_UNKNOWN_HANDLE, "unknown row set %llu", (unsigned long long)plan_->scan_row_set);
      if (plan_->nterms || plan_->nbloom) return not_eligible("row-set scans carry no predicate terms or Bloom probes");
      const JoinTable& rs = rt->second;
      rowset_ = &rs;
      L_->rowscan = true;
      L_->row_scan.schema.push_back(pgf_column_spec{uint16_t(rs.key_type), 0});
      for (uint32_t p = 0; p < rs.npayload; ++p) L_->row_scan.schema.push_back(pgf_column_spec{uint16_t(rs.payload_type[p]), rs.payload_nullable[p]});
      L_->row_scan.rows = rs.rows;
      L_->row_scan.finished = true;
      L_->dev.row_src = rs.d_rows;
      L_->dev.row_count = rs.rows;
      L_->dev.row_u4 = rs.slot_u4;
    }
    auto it = ctx_->scans.find(plan_->scan_id);
    if (!L_->rowscan && it == ctx_->scans.end()) return ctx_->fail(PGF_ERR_UNKNOWN_HANDLE, "unknown scan %llu", (unsigned long long)plan_->scan_id);
    Scan& s = L_->rowscan ? L_->row_scan : *it->second;
    if (!s.finished) return ctx_->fail(PGF_ERR_STATE, "scan %llu is not finished", (unsigned long long)plan_->scan_id);
    L_->scan = &s;
    DevPlan& D = L_->dev;
    if (plan_->nbloom > kMaxBlooms || plan_->nterms > kMaxTerms || plan_->njoins > kMaxJoins ||
        plan_->nkeys > 4 || plan_->nexprs > kMaxExprs || plan_->naggs > PGF_MAX_AGGS || plan_->npayload > PGF_MAX_PAYLOAD)
      return not_eligible("plan exceeds the fixed operator limits");
    // Pipelines with a join probe or a build sink run the compaction kernel: only the predicate
    // columns, the Bloom keys and the first probe key are staged; everything a surviving row needs
    // later is read from the page in HBM (late refs).
    probe_ = plan_->njoins > 0 || plan_->sink == PGF_SINK_JOIN_BUILD || L_->rowscan;
    L_->probe = probe_;

    // joins first: payload refs need the tables
    for (uint32_t j = 0; j < plan_->njoins; ++j) {
      auto jt = ctx_->joins.find(plan_->joins[j].join_table);
      if (jt == ctx_->joins.end()) return ctx_->fail(PGF_ERR_UNKNOWN_HANDLE, "unknown join table %llu", (unsigned long long)plan_->joins[j].join_table);
      jtables_[j] = &jt->second;
      if (!jt->second.d_slots) return ctx_->fail(PGF_ERR_STATE, "join input %llu is a row set: it has no hash table to probe", (unsigned long long)plan_->joins[j].join_table);
    }
    for (uint32_t j = 0; j < plan_->njoins; ++j) {
      DevJoin& dj = D.joins[j];
      dj.slots = jtables_[j]->d_slots;
      dj.tags = reinterpret_cast<const uint8_t*>(jtables_[j]->d_slots) + uint64_t(jtables_[j]->capacity) * jtables_[j]->slot_u4 * sizeof(uint4);
      dj.mask = jtables_[j]->capacity - 1;
      dj.slot_u4 = jtables_[j]->slot_u4;
      dj.shift = 64;
      for (uint32_t c = jtables_[j]->capacity; c > 1; c >>= 1) dj.shift--;
      PGF_TRY(lower_ref(plan_->joins[j].probe_key, j, &dj.key, /*late=*/j > 0 || L_->rowscan));
      if (!is_int_type(dj.key.type)) return not_eligible("join keys must be Int16/Int32/Int64");
    }
    D.njoins = plan_->njoins;
    L_->nj = plan_->njoins;


    for (uint32_t b = 0; b < plan_->nbloom; ++b) {
      auto bt = ctx_->blooms.find(plan_->bloom[b].bloom);
      if (bt == ctx_->blooms.end()) return ctx_->fail(PGF_ERR_UNKNOWN_HANDLE, "unknown bloom filter %llu", (unsigned long long)plan_->bloom[b].bloom);
      // Only a Ready filter of the expected generation may reject rows; anything else is
      // PassUnfiltered (runtime_filter/src/shared.rs:350-361), i.e. the probe is dropped.
      const BloomSlot& bs = bt->second;
      if ((bs.lifecycle >> 2) != plan_->bloom[b].expected_generation || int(bs.lifecycle & 3) != PGF_RF_READY) {
        L_->bloom_dropped++;
        continue;
      }
      DevBloomProbe& bp = D.bloom[D.nbloom];
      bp.bloom = bs.dev;
      PGF_TRY(lower_ref(plan_->bloom[b].key, 0, &bp.key));
      if (bp.key.src != SRC_PAGE) return not_eligible("Bloom probe keys must be scan columns");
      if (!is_int_type(bp.key.type)) return not_eligible("runtime filter keys must be Int16/Int32/Int64");
      if (!(ctx_->flags & PGF_CFG_KEEP_REDUNDANT_BLOOM_PROBES)) {
        // (a) the same pipeline probes a join table on this key: its tag directory rejects misses just as cheaply
        bool redundant = false;
        for (uint32_t j = 0; j < D.njoins; ++j)
          redundant |= D.joins[j].key.src == SRC_PAGE && D.joins[j].key.pcol == bp.key.pcol;
        // (b) saturated filter: an absent key passes with probability fill^k; above 0.9 the probe costs more than it rejects
        const double fill = double(bs.set_bits) / double(bs.params.bit_count);
        redundant |= std::pow(fill, double(bs.params.hash_count)) > 0.9;
        if (redundant) {   // PassUnfiltered for every row: the result is the same
          L_->bloom_dropped++;
          continue;
        }
      }
      D.nbloom++;
    }

    PGF_TRY(lower_terms());

    switch (plan_->sink) {
      case PGF_SINK_AGGREGATE: PGF_TRY(lower_aggregate()); break;
      case PGF_SINK_JOIN_BUILD: PGF_TRY(lower_join_build()); break;
      case PGF_SINK_COUNT: D.sink = SINK_COUNT; break;
      default: return ctx_->fail(PGF_ERR_INVALID_ARGUMENT, "unknown sink %d", plan_->sink);
    }
    // common subexpression: x*(c-y)*(c2+z) right after x*(c-y) reuses the previous value
    // (ref.off still holds the stage slot of the column here: equal slots <=> equal columns)
    for (uint32_t e = 1; e < D.nexprs && !probe_; ++e) {
      const DevExpr &a = D.exprs[e - 1], &b = D.exprs[e];
      if (a.form == FORM_X_CMY && b.form == FORM_X_CMY_CPZ && a.f[0].ref.off == b.f[0].ref.off &&
          a.f[1].ref.off == b.f[1].ref.off && a.f[1].cf == b.f[1].cf && a.f[1].ci_lo == b.f[1].ci_lo && a.f[1].ci_hi == b.f[1].ci_hi)
        D.exprs[e].form = FORM_PREV_CPZ;
    }
    // The order of the conjuncts is irrelevant to the result: if some order of the range terms
    // matches a registered shape, take it.
    if (!probe_ && !pick_shape(*L_) && D.nterms >= 2 && D.nterms <= 4) {
      bool plain = true;
      for (uint32_t t = 0; t < D.nterms; ++t) plain &= D.terms[t].op == TERM_IN_RANGE;
      if (plain) {
        DevTerm orig[4];
        uint32_t perm[4] = {0, 1, 2, 3};
        for (uint32_t t = 0; t < D.nterms; ++t) orig[t] = D.terms[t];
        bool found = false;
        while (!found && std::next_permutation(perm, perm + D.nterms)) {
          for (uint32_t t = 0; t < D.nterms; ++t) D.terms[t] = orig[perm[t]];
          found = pick_shape(*L_) != nullptr;
        }
        if (!found)
          for (uint32_t t = 0; t < D.nterms; ++t) D.terms[t] = orig[t];
      }
    }
    if (probe_ && !L_->rowscan && D.nbloom) {
      // A runtime filter keyed on the column the compacted entries carry is probed after the predicate, on dense
      // lanes: a Bloom probe costs ~120 instructions, so it should only see rows the predicate kept.  Put it first.
      const int want = D.njoins ? int(D.joins[0].key.pcol) : -1;
      for (uint32_t b = 0; b < D.nbloom; ++b) {
        if (want < 0 || int(D.bloom[b].key.pcol) == want) {
          std::swap(D.bloom[0], D.bloom[b]);
          D.bloom_dense = 1;
          break;
        }
      }
    }
    if (L_->rowscan) {
      D.nitems = s.rows ? 1 : 0;
    } else if (probe_) {
      PGF_TRY(layout_stage_probe(s));
    } else {
      PGF_TRY(layout_stage(s));
    }
    fix_refs();
    if (probe_ && !L_->rowscan && (D.njoins || D.bloom_dense) && D.nbloom == D.bloom_dense) {
      // the specialisation of probe_pipeline_kernel: one plain string range, nothing nullable, Int32 entry key
      const DevRef& kr = D.njoins ? D.joins[0].key : D.bloom[0].key;
      if (D.nterms == 1 && D.terms[0].op == TERM_IN_RANGE && D.used_null_mask == 0 && D.terms[0].ref.ld == LD_VIEW &&
          kr.src == SRC_PAGE && kr.ld == LD_I32) {
        L_->t0 = LD_VIEW;
        D.entry_key_off = kr.off;
      }
    }
    return PGF_OK;
  }

 private:
  // FilterExec conjuncts -> one inclusive key range per column (+ one term per NE).
  pgf_status lower_terms() {
    DevPlan& D = L_->dev;
    struct Range {
      pgf_colref col;
      DevRef ref;
      bool wide;
      __int128 lo, hi;  // over the 128-bit key ((i128)k0 << 64 | k1)
    };
    std::vector<Range> ranges;
    const __int128 kMin = (__int128)1 << 127, kMax = ~kMin;
    for (uint32_t t = 0; t < plan_->nterms; ++t) {
      const pgf_pred_term& pt = plan_->terms[t];
      if (pt.col.source != 0) return not_eligible("predicates are evaluated on scan columns");
      if (pt.cmp < PGF_CMP_LT || pt.cmp > PGF_CMP_NE) return ctx_->fail(PGF_ERR_INVALID_ARGUMENT, "bad comparison operator");
      DevRef ref;
      PGF_TRY(lower_ref(pt.col, plan_->njoins, &ref, /*late=*/false, /*predicate=*/true));
      int64_t k0 = 0;
      uint64_t k1 = 0;
      PGF_TRY(literal_key(pt.lit, ref.type, &k0, &k1));
      const bool wide = ref.ld == LD_VIEW || ref.ld == LD_DEC;
      const __int128 key = ((__int128)k0 << 64) | (__int128)k1;
      // one-word keys only occupy k1 == 0: step over whole words so that "< c" is "<= c - 1"
      const __int128 one = wide ? (__int128)1 : ((__int128)1 << 64);
      if (pt.cmp == PGF_CMP_NE) {
        if (D.nterms >= kMaxTerms) return not_eligible("too many predicate terms");
        DevTerm& dt = D.terms[D.nterms++];
        dt.ref = ref;
        dt.op = TERM_NOT_IN_RANGE;
        set_bounds(&dt, wide, key, key);
        continue;
      }
      Range* rg = nullptr;
      for (auto& r : ranges)
        if (r.col.col == pt.col.col) rg = &r;
      if (!rg) {
        ranges.push_back(Range{pt.col, ref, wide, kMin, kMax});
        rg = &ranges.back();
      }
      switch (pt.cmp) {
        case PGF_CMP_LT: if (key == kMin) rg->hi = kMin, rg->lo = kMax; else rg->hi = std::min(rg->hi, key - one); break;
        case PGF_CMP_LE: rg->hi = std::min(rg->hi, key); break;
        case PGF_CMP_GT: if (key > kMax - one) rg->hi = kMin, rg->lo = kMax; else rg->lo = std::max(rg->lo, key + one); break;
        case PGF_CMP_GE: rg->lo = std::max(rg->lo, key); break;
        default: rg->lo = std::max(rg->lo, key); rg->hi = std::min(rg->hi, key); break;  // EQ
      }
    }
    for (auto& r : ranges) {
      if (D.nterms >= kMaxTerms) return not_eligible("too many predicate terms");
      DevTerm& dt = D.terms[D.nterms++];
      dt.ref = r.ref;
      if (r.lo > r.hi) {
        dt.op = TERM_NEVER;
        set_bounds(&dt, r.wide, 0, 0);
      } else {
        dt.op = TERM_IN_RANGE;
        set_bounds(&dt, r.wide, r.lo, r.hi);
      }
    }
    return PGF_OK;
  }

  // Narrow keys: lo0 and span.  Wide keys: the device compares unsigned 128-bit values, so the
  // signed order of (k0, k1) is mapped by flipping the sign bit; stored as lo and span = hi - lo.
  static void set_bounds(DevTerm* dt, bool wide, __int128 lo, __int128 hi) {
    dt->wide = wide;
    if (wide) {
      const unsigned __int128 flip = (unsigned __int128)1 << 127;
      const unsigned __int128 ulo = (unsigned __int128)lo ^ flip, uhi = (unsigned __int128)hi ^ flip;
      const unsigned __int128 span = uhi - ulo;
      dt->lo0 = int64_t(uint64_t(ulo >> 64));
      dt->lo1 = uint64_t(ulo);
      dt->hi0 = int64_t(uint64_t(span >> 64));
      dt->hi1 = uint64_t(span);
    } else {
      dt->lo0 = int64_t(lo >> 64);
      dt->hi0 = int64_t(hi >> 64);
      dt->lo1 = uint64_t(dt->hi0) - uint64_t(dt->lo0);  // span
      dt->hi1 = 0;
    }
  }

  pgf_status not_eligible(const char* why) { return ctx_->fail(PGF_ERR_NOT_ELIGIBLE, "pipeline not eligible for the GPU path: %s", why); }

  static uint8_t ld_kind(int type) {
    switch (type) {
      case PGF_T_INT16: return LD_I16;
      case PGF_T_INT32: return LD_I32;
      case PGF_T_INT64: return LD_I64;
      case PGF_T_FLOAT32: return LD_F32;
      case PGF_T_FLOAT64: return LD_F64;
      case PGF_T_DECIMAL128: return LD_DEC;
      case PGF_T_BOOLEAN: return LD_BOOL;
      default: return LD_VIEW;
    }
  }

  // Resolve a column reference; scan columns get a stage slot.  For scan columns `off`
  // temporarily holds the slot; fix_ref() turns it into shared-memory offsets once the
  // stage layout is known.
  static constexpr uint32_t kLateRef = 0xFFFFFFFEu;  // `off` of a page column that is not staged (read from HBM by stage C)
  pgf_status lower_ref(const pgf_colref& r, uint32_t joins_visible, DevRef* out, bool late = false, bool predicate = false) {
    Scan& s = *L_->scan;
    DevPlan& D = L_->dev;
    if (r.source == 0) {
      if (r.col < 0 || size_t(r.col) >= s.schema.size()) return ctx_->fail(PGF_ERR_INVALID_ARGUMENT, "column %d out of range", r.col);
      const int type = s.schema[r.col].type_tag;
      // Boolean columns (bit-packed values) are evaluated in predicates only; Uuid columns nowhere
      if (type == PGF_T_UUID || (type == PGF_T_BOOLEAN && (!predicate || L_->rowscan)))
        return not_eligible("Uuid columns, and Boolean columns outside predicates, are not evaluated on the GPU path");
      if (L_->rowscan) {  // a column of the scanned row: word 0..1 = key, 3 + payload word = payload (GRow::rec)
        out->src = kSrcRecord;
        out->ld = ld_kind(type);
        out->type = uint8_t(type);
        out->pcol = uint8_t(r.col == 0 ? 30 : r.col - 1);   // bit 1 + pcol of the row's flag word says NULL; the key never is
        out->off = r.col == 0 ? 0u : 3u + rowset_->payload_word[r.col - 1];
        out->valid_off = kNoValidity;
        return PGF_OK;
      }
      if (late && probe_) {
        bool listed = false;
        for (uint32_t c = 0; c < D.nlate; ++c) listed |= D.late_pcol[c] == uint8_t(r.col);
        if (!listed && D.nlate < 8) {
          D.late_pcol[D.nlate] = uint8_t(r.col);
          D.late_width[D.nlate++] = uint8_t(row_width(type));
        }
        out->src = SRC_PAGE;
        out->ld = ld_kind(type);
        out->type = uint8_t(type);
        out->pcol = uint8_t(r.col);
        out->off = kLateRef;
        out->valid_off = s.schema[r.col].nullable ? 0u : kNoValidity;
        if (s.schema[r.col].nullable) D.used_null_mask |= 1u << r.col;
        if (is_view(type)) D.view_mask |= 1u << r.col;
        return PGF_OK;
      }
      uint32_t slot = 0;
      for (; slot < D.nstage_cols; ++slot)
        if (D.scol[slot].page_col == r.col) break;
      if (slot == D.nstage_cols) {
        DevStageCol& sc = D.scol[D.nstage_cols++];
        sc.page_col = uint16_t(r.col);
        sc.width = uint16_t(row_width(type));
        sc.nullable = s.schema[r.col].nullable;
        sc.type = uint16_t(type);
        if (sc.nullable) D.used_null_mask |= 1u << r.col;
        if (is_view(type)) D.view_mask |= 1u << r.col;
      }
      out->src = SRC_PAGE;
      out->ld = ld_kind(type);
      out->type = uint8_t(type);
      out->pcol = uint8_t(r.col);
      out->off = slot;
      out->valid_off = kNoValidity;
      return PGF_OK;
    }
    const uint32_t j = uint32_t(r.source - 1);
    if (r.source < 0 || j >= joins_visible) return ctx_->fail(PGF_ERR_INVALID_ARGUMENT, "column source %d is not visible here", r.source);
    const JoinTable& jt = *jtables_[j];
    if (r.col < 0 || uint32_t(r.col) >= jt.npayload) return ctx_->fail(PGF_ERR_INVALID_ARGUMENT, "join payload %d out of range", r.col);
    out->src = uint8_t(j + 1);
    out->ld = ld_kind(jt.payload_type[r.col]);
    out->type = uint8_t(jt.payload_type[r.col]);
    out->pcol = uint8_t(r.col);
    out->off = jt.payload_word[r.col];
    out->valid_off = kNoValidity;
    return PGF_OK;
  }

  // Is the value behind a column reference declared NOT NULL?  (scan column, or build-side column of a join)
  bool ref_not_null(const pgf_colref& r) const {
    if (r.source == 0) return !L_->scan->schema[r.col].nullable;
    return !jtables_[r.source - 1]->payload_nullable[r.col];
  }

  void fix_ref(DevRef* ref) {
    if (ref->src != SRC_PAGE) return;
    if (ref->off == kLateRef) { ref->off = 0; return; }
    const DevStageCol& sc = L_->dev.scol[ref->off];
    ref->valid_off = sc.nullable ? sc.valid_off : kNoValidity;
    ref->off = sc.smem_off;
  }

  // Second pass once layout_stage() has placed the column tiles.
  void fix_refs() {
    DevPlan& D = L_->dev;
    for (uint32_t b = 0; b < D.nbloom; ++b) fix_ref(&D.bloom[b].key);
    for (uint32_t t = 0; t < D.nterms; ++t) fix_ref(&D.terms[t].ref);
    for (uint32_t j = 0; j < D.njoins; ++j) fix_ref(&D.joins[j].key);
    for (uint32_t k = 0; k < D.nkeys; ++k) fix_ref(&D.keys[k].ref);
    for (uint32_t e = 0; e < D.nexprs; ++e)
      for (uint32_t f = 0; f < D.exprs[e].nfactors; ++f) fix_ref(&D.exprs[e].f[f].ref);
    fix_ref(&D.build.key);
    for (uint32_t p = 0; p < D.build.npayload; ++p) fix_ref(&D.build.payload[p]);
  }

  // Literal -> order-preserving key, matching term_pass2 on the device.
  pgf_status literal_key(const pgf_literal& lit, int col_type, int64_t* k0, uint64_t* k1) {
    *k1 = 0;
    switch (col_type) {
      case PGF_T_BOOLEAN:
        if (lit.type_tag != PGF_T_BOOLEAN || (lit.i64 != 0 && lit.i64 != 1)) return not_eligible("Boolean column compared with a non-Boolean literal");
        *k0 = lit.i64;   // false < true, as in arrow's comparison kernels
        return PGF_OK;
      case PGF_T_INT16: case PGF_T_INT32: case PGF_T_INT64:
        if (lit.type_tag != PGF_T_INT64 && lit.type_tag != PGF_T_INT32 && lit.type_tag != PGF_T_INT16)
          return not_eligible("integer column compared with a non-integer literal");
        *k0 = lit.i64;
        return PGF_OK;
      case PGF_T_FLOAT64: {
        double d;
        if (lit.type_tag == PGF_T_FLOAT64) d = lit.f64;
        else if (lit.type_tag == PGF_T_INT64) d = double(lit.i64);
        else return not_eligible("Float64 column compared with an unsupported literal");
        int64_t b;
        std::memcpy(&b, &d, 8);
        b ^= int64_t(uint64_t(b >> 63) >> 1);  // IEEE totalOrder key
        *k0 = b;
        return PGF_OK;
      }
      case PGF_T_DECIMAL128:
        if (lit.type_tag != PGF_T_DECIMAL128) return not_eligible("Decimal128 column compared with a non-decimal literal");
        *k0 = lit.hi;
        *k1 = uint64_t(lit.i64);
        return PGF_OK;
      case PGF_T_UTF8VIEW: case PGF_T_BINARYVIEW: {
        if (lit.type_tag != PGF_T_UTF8VIEW && lit.type_tag != PGF_T_BINARYVIEW) return not_eligible("string column compared with a non-string literal");
        if (lit.slen < 0 || lit.slen > 12) return not_eligible("string literals longer than 12 bytes need the out-of-line view path");
        uint8_t pad[12] = {0};
        std::memcpy(pad, lit.str, size_t(lit.slen));
        uint64_t hi = 0;
        for (int i = 0; i < 8; ++i) hi = (hi << 8) | pad[i];
        uint32_t lo = 0;
        for (int i = 8; i < 12; ++i) lo = (lo << 8) | pad[i];
        *k0 = int64_t(hi ^ 0x8000000000000000ull);
        *k1 = (uint64_t(lo) << 32) | uint32_t(lit.slen);
        return PGF_OK;
      }
      default:
        return not_eligible("predicate on a column type the GPU path does not compare");
    }
  }

  pgf_status lower_aggregate() {
    DevPlan& D = L_->dev;
    D.sink = SINK_AGG;
    if (plan_->naggs == 0 && plan_->nkeys == 0) return ctx_->fail(PGF_ERR_INVALID_ARGUMENT, "aggregate without keys or aggregates");
    // accumulator class: all expressions of one pipeline share it
    int cls = -1;
    for (uint32_t e = 0; e < plan_->nexprs; ++e) {
      const pgf_value_expr& x = plan_->exprs[e];
      if (x.nfactors < 1 || x.nfactors > 3) return ctx_->fail(PGF_ERR_INVALID_ARGUMENT, "expression %u has %u factors", e, x.nfactors);
      DevExpr& dx = D.exprs[e];
      dx.nfactors = x.nfactors;
      int ecls = -1;
      bool all_i64 = true;
      for (uint32_t f = 0; f < x.nfactors; ++f) {
        DevFactor& df = dx.f[f];
        PGF_TRY(lower_ref(x.factors[f].col, plan_->njoins, &df.ref, /*late=*/true));
        df.kind = uint32_t(x.factors[f].kind);
        if (df.kind > PGF_FACTOR_CONST_PLUS_COL) return ctx_->fail(PGF_ERR_INVALID_ARGUMENT, "bad factor kind");
        int fcls;
        if (df.ref.type == PGF_T_FLOAT64) fcls = CLS_F64;
        else if (df.ref.type == PGF_T_DECIMAL128) fcls = CLS_I128;
        else if (is_int_type(df.ref.type)) fcls = CLS_I64;
        else return not_eligible("aggregate argument of a type the GPU path does not compute");
        all_i64 &= df.ref.type == PGF_T_INT64;
        if (ecls >= 0 && ecls != fcls) return not_eligible("mixed-type arithmetic in one expression");
        ecls = fcls;
        const pgf_literal& c = x.factors[f].c;
        if (df.kind != PGF_FACTOR_COL) {
          if (fcls == CLS_F64) df.cf = c.type_tag == PGF_T_FLOAT64 ? c.f64 : double(c.i64);
          else { df.ci_lo = c.i64; df.ci_hi = c.type_tag == PGF_T_DECIMAL128 ? c.hi : (c.i64 < 0 ? -1 : 0); }
        }
      }
      dx.form = FORM_GENERIC;
      dx.null_cols = 0;
      dx.has_payload = 0;
      bool plain_f64 = true, plain_dec = true;
      for (uint32_t f = 0; f < x.nfactors; ++f) {
        const DevRef& rf = dx.f[f].ref;
        plain_f64 &= rf.src == SRC_PAGE && rf.ld == LD_F64;
        plain_dec &= rf.src == SRC_PAGE && rf.ld == LD_DEC;
        if (rf.src == SRC_PAGE) { if (L_->scan->schema[rf.pcol].nullable) dx.null_cols |= 1u << rf.pcol; }
        else dx.has_payload = 1;
      }
      if (plain_f64 || plain_dec) {  // straight-line forms (the generic evaluators ignore them)
        const uint32_t k0 = dx.f[0].kind, k1 = x.nfactors > 1 ? dx.f[1].kind : 0, k2 = x.nfactors > 2 ? dx.f[2].kind : 0;
        if (x.nfactors == 1 && k0 == PGF_FACTOR_COL) dx.form = FORM_X;
        else if (x.nfactors == 2 && k0 == PGF_FACTOR_COL && k1 == PGF_FACTOR_COL) dx.form = FORM_XY;
        else if (x.nfactors == 2 && k0 == PGF_FACTOR_COL && k1 == PGF_FACTOR_CONST_MINUS_COL) dx.form = FORM_X_CMY;
        else if (x.nfactors == 3 && k0 == PGF_FACTOR_COL && k1 == PGF_FACTOR_CONST_MINUS_COL && k2 == PGF_FACTOR_CONST_PLUS_COL) dx.form = FORM_X_CMY_CPZ;
      }
      // narrow integers wrap at their own width in arrow; only plain columns or all-Int64
      // arithmetic is bit-exact with 64-bit registers
      if (ecls == CLS_I64 && !(all_i64 || (x.nfactors == 1 && dx.f[0].kind == PGF_FACTOR_COL)))
        return not_eligible("arithmetic on Int16/Int32 columns");
      expr_cls_[e] = ecls;
    }
    // AVG over integers is computed on the Float64 cast (DataFusion coerces avg(int) to Float64)
    for (uint32_t a = 0; a < plan_->naggs; ++a) {
      const pgf_agg& ag = plan_->aggs[a];
      if (ag.func < PGF_AGG_SUM || ag.func > PGF_AGG_COUNT) return ctx_->fail(PGF_ERR_INVALID_ARGUMENT, "bad aggregate function");
      if (ag.func == PGF_AGG_COUNT_STAR) continue;
      if (ag.expr < 0 || uint32_t(ag.expr) >= plan_->nexprs) return ctx_->fail(PGF_ERR_INVALID_ARGUMENT, "aggregate %u references expression %d", a, ag.expr);
      if (ag.func == PGF_AGG_AVG && expr_cls_[ag.expr] == CLS_I64) expr_as_f64_[ag.expr] = true;
    }
    for (uint32_t e = 0; e < plan_->nexprs; ++e) {
      int ecls = expr_cls_[e];
      if (expr_as_f64_[e]) {
        if (plan_->exprs[e].nfactors != 1 || plan_->exprs[e].factors[0].kind != PGF_FACTOR_COL) return not_eligible("AVG over integer arithmetic");
        for (uint32_t a = 0; a < plan_->naggs; ++a)
          if (plan_->aggs[a].func != PGF_AGG_COUNT_STAR && plan_->aggs[a].expr == int32_t(e) && plan_->aggs[a].func == PGF_AGG_SUM)
            return not_eligible("SUM and AVG sharing one integer expression");
        ecls = CLS_F64;
      }
      if (cls >= 0 && cls != ecls) return not_eligible("aggregates of different accumulator classes in one pipeline");
      cls = ecls;
    }
    if (cls < 0) cls = CLS_I64;  // COUNT(*) only
    const uint32_t max_exprs = cls == CLS_I128 ? kAccI128MaxExprs : kAccF64MaxExprs;
    if (plan_->nexprs > max_exprs) return not_eligible("too many distinct aggregate arguments");
    D.nexprs = plan_->nexprs;
    D.acc_cls = uint32_t(cls);
    L_->acc_cls = uint32_t(cls);
    L_->maxe = plan_->nexprs <= 2 ? 2 : max_exprs;
    // an integer expression averaged as Float64 must not take the raw-f64 fast forms
    for (uint32_t e = 0; e < plan_->nexprs; ++e)
      if (expr_as_f64_[e]) D.exprs[e].form = FORM_GENERIC;
    // Canonical argument order on the device: by form (x, x*y, x*(c-y), x*(c-y)*(c+z), generic), stable.
    // The order aggregates are written in then no longer decides whether a registered shape matches,
    // and x*(c-y)*(c+z) lands right behind its x*(c-y) for the common-subexpression rewrite.
    {
      auto rank = [](uint32_t form) { return form == FORM_GENERIC ? 99u : form; };
      uint32_t order[kMaxExprs];
      for (uint32_t e = 0; e < plan_->nexprs; ++e) order[e] = e;
      std::stable_sort(order, order + plan_->nexprs, [&](uint32_t a, uint32_t b) { return rank(D.exprs[a].form) < rank(D.exprs[b].form); });
      DevExpr tmp[kMaxExprs];
      for (uint32_t e = 0; e < plan_->nexprs; ++e) tmp[e] = D.exprs[e];
      for (uint32_t pos = 0; pos < plan_->nexprs; ++pos) {
        D.exprs[pos] = tmp[order[pos]];
        L_->expr_pos[order[pos]] = pos;
      }
    }

    uint32_t words = 0;
    for (uint32_t k = 0; k < plan_->nkeys; ++k) {
      DevKeyPart& kp = D.keys[k];
      PGF_TRY(lower_ref(plan_->keys[k], plan_->njoins, &kp.ref, /*late=*/true));
      const int t = kp.ref.type;
      if (is_int_type(t)) kp.nwords = 1;
      else if (is_view(t) || t == PGF_T_DECIMAL128) kp.nwords = 2;
      else return not_eligible("group key of a type the GPU path does not hash");
      kp.word = uint16_t(words);
      words += kp.nwords;
      L_->key_types[k] = t;
      L_->key_not_null[k] = ref_not_null(plan_->keys[k]);
    }
    if (words > kKeyWords) return not_eligible("group key wider than 32 bytes");
    D.nkeys = plan_->nkeys;
    D.nkeywords = words;
    L_->grouped = plan_->nkeys > 0;
    if (L_->grouped) {
      uint64_t want = plan_->expected_groups ? plan_->expected_groups * 2 : (1ull << 16);
      uint64_t cap = 1024;
      while (cap < want) cap <<= 1;
      if (cap > (1ull << 30)) cap = 1ull << 30;
      L_->table_capacity = cap;
    } else {
      L_->table_capacity = 1;
    }
    return PGF_OK;
  }

  pgf_status lower_join_build() {
    DevPlan& D = L_->dev;
    D.sink = SINK_JOIN_BUILD;
    JoinBuild& jb = D.build;
    PGF_TRY(lower_ref(plan_->build_key, plan_->njoins, &jb.key, /*late=*/true));
    if (!is_int_type(jb.key.type)) return not_eligible("join keys must be Int16/Int32/Int64");
    JoinTable& jt = L_->build_table;
    jt.key_type = jb.key.type;
    jt.npayload = plan_->npayload;
    uint32_t words = 0;
    for (uint32_t p = 0; p < plan_->npayload; ++p) {
      PGF_TRY(lower_ref(plan_->payload[p], plan_->njoins, &jb.payload[p], /*late=*/true));
      const uint32_t w = type_u32_words(jb.payload[p].type);
      if (!w) return not_eligible("join payload column type");
      jb.payload_word[p] = uint16_t(words);
      jb.payload_nwords[p] = uint16_t(w);
      jt.payload_type[p] = jb.payload[p].type;
      jt.payload_nullable[p] = ref_not_null(plan_->payload[p]) ? 0 : 1;
      jt.payload_word[p] = uint16_t(words);
      words += w;
    }
    if (words > 5) return not_eligible("join payload wider than 20 bytes");
    jb.npayload = plan_->npayload;
    jb.slot_u4 = words > 1 ? 2 : 1;
    jt.slot_u4 = jb.slot_u4;
    if (plan_->build_bloom) {
      auto bt = ctx_->blooms.find(plan_->build_bloom);
      if (bt == ctx_->blooms.end()) return ctx_->fail(PGF_ERR_UNKNOWN_HANDLE, "unknown bloom filter %llu", (unsigned long long)plan_->build_bloom);
      if (int(bt->second.lifecycle & 3) != PGF_RF_BUILDING)
        return ctx_->fail(PGF_ERR_LIFECYCLE_INVALID_TRANSITION, "build_bloom must be in Building state");
      D.build_bloom = bt->second.dev;
      D.has_build_bloom = 1;
    }
    return PGF_OK;
  }

  // Shared-memory stage layout: row tiles of at most ~40 KiB, tile_rows a multiple of 128 so
  // every column slice (and validity slice) starts 16-byte aligned.
  static uint32_t col_tile_bytes(const DevStageCol& sc, uint32_t tile_rows) { return sc.width ? tile_rows * sc.width : tile_rows / 8u; }

  pgf_status layout_stage(Scan& s) {
    DevPlan& D = L_->dev;
    uint32_t row_bytes8 = 0;  // bytes per row x 8 (validity counts 1 bit)
    for (uint32_t c = 0; c < D.nstage_cols; ++c) row_bytes8 += (D.scol[c].width ? D.scol[c].width * 8u : 1u) + (D.scol[c].nullable ? 1u : 0u);
    const uint32_t max_rows = s.max_page_rows ? s.max_page_rows : 1;
    const uint64_t page_bytes = (uint64_t(row_bytes8) * max_rows + 7) / 8;
    // Ring shape.  Streaming pipelines: kStages tiles of ~40 KiB (a fraction of a page).  Behind a
    // join probe a thread carries 4 rows, so a tile must hold consumer_warps x 128 rows to keep
    // every warp busy: whole pages in a 3-deep ring.  (Measured on Q1: whole-page tiles raise the
    // share of busy warp slots from 46 % to 93 % but do not help -- the kernel is issue bound, and
    // the deeper ring of smaller tiles hides the TMA latency better.)
    uint32_t queue_bytes = (D.sink == SINK_AGG && L_->grouped) ? kMaxConsumerWarps * kQueueBytesPerWarp : 0u;
    // The fast GROUP BY path sends the rows it cannot keep in its slots straight to the global table: it has no
    // deferred-sink queues, and their 32 KiB go to the ring instead (Q1: 4 x 416-row tiles instead of 4 x 288 --
    // fewer tile hand-overs per row and 100 KB instead of 69 KB in flight per SM).
    if (const ShapeEntry* se = pick_shape(*L_))
      if (se->fast_group_exprs) queue_bytes = fast_group_acc_bytes(se->fast_group_exprs);  // accumulator slots
    const uint32_t smem_budget = 227u * 1024u - uint32_t((sizeof(BlockShared) + 127) & ~size_t(127)) - queue_bytes;
    // tile starts must be 16-byte aligned in every staged buffer: 128 rows when a validity bitmap
    // is staged, else 16 rows (2-byte values)
    uint32_t gran = 16;
    for (uint32_t c = 0; c < D.nstage_cols; ++c)
      if (D.scol[c].nullable || D.scol[c].width == 0) gran = 128;   // (a Boolean column is a bitmap itself)
    auto shape = [&](uint32_t ntiles, uint32_t* tile_rows_out, uint32_t* stage_bytes_out) {
      const uint32_t tile_rows = ((max_rows + ntiles - 1) / ntiles + gran - 1) / gran * gran;
      uint32_t stage_bytes = 0;
      for (uint32_t c = 0; c < D.nstage_cols; ++c) stage_bytes += col_tile_bytes(D.scol[c], tile_rows);
      for (uint32_t c = 0; c < D.nstage_cols; ++c)
        if (D.scol[c].nullable) stage_bytes += tile_rows / 8;
      *tile_rows_out = tile_rows;
      *stage_bytes_out = (stage_bytes + 127u) & ~127u;
    };
    uint32_t best_tiles = 0, best_stages = 0;
    {
      uint32_t tile_rows, stage_bytes;
      shape(1, &tile_rows, &stage_bytes);
      if (L_->nj != 0 && stage_bytes <= ((smem_budget / 3u) & ~127u)) {
        best_tiles = 1;
        best_stages = 3;
      } else {
        const uint32_t cap = std::min<uint32_t>((smem_budget / kStages) & ~127u, 48u * 1024u);
        uint32_t ntiles = uint32_t((page_bytes + 40u * 1024u - 1) / (40u * 1024u));
        for (ntiles = ntiles ? ntiles : 1;; ++ntiles) {
          shape(ntiles, &tile_rows, &stage_bytes);
          if (stage_bytes <= cap) { best_tiles = ntiles; best_stages = kStages; break; }
          if (tile_rows <= gran) break;
        }
      }
    }
    if (!best_tiles) return not_eligible("row too wide for the shared-memory stages");
    uint32_t tile_rows = 0, stage_bytes = 0;
    shape(best_tiles, &tile_rows, &stage_bytes);
    {
      uint32_t off = 0;
      for (uint32_t c = 0; c < D.nstage_cols; ++c) {
        D.scol[c].smem_off = off;
        off += col_tile_bytes(D.scol[c], tile_rows);
      }
      for (uint32_t c = 0; c < D.nstage_cols; ++c) {
        D.scol[c].valid_off = off;
        if (D.scol[c].nullable) off += tile_rows / 8;
      }
    }
    const uint32_t nstages = best_stages;
    D.nstages = nstages;
    D.tile_rows = tile_rows;
    D.tiles_per_page = (max_rows + tile_rows - 1) / tile_rows;
    D.stage_bytes = stage_bytes ? stage_bytes : 128;
    D.npages = uint32_t(s.npages);
    if (s.npages * uint64_t(D.tiles_per_page) > 0xFFFFFFF0ull) return not_eligible("scan too large for 32-bit tile ids");
    D.nitems = uint32_t(s.npages) * D.tiles_per_page;
    D.pages = s.d_pages;
    D.descs = s.d_descs;
    D.classes = s.d_classes;
    D.page_stride = ctx_->page_size;
    L_->smem = ((sizeof(BlockShared) + 127) & ~size_t(127)) + size_t(nstages) * D.stage_bytes + queue_bytes;  // ring + deferred-sink queues
    return PGF_OK;
  }

  // Per-warp tile rings of the compaction pipeline: tiles of 128 rows (fewer when the staged row is
  // wide), as deep as the warp's share of shared memory allows (2 .. kPMaxDepth tiles in flight).
  pgf_status layout_stage_probe(Scan& s) {
    DevPlan& D = L_->dev;
    const uint32_t max_rows = s.max_page_rows ? s.max_page_rows : 1;
    const uint32_t budget = ((227u * 1024u - probe_shared_bytes()) / uint32_t(kPConsumerWarps) - kPQueueBytesPerWarp) & ~127u;
    bool nullable = false;   // some staged buffer is a bitmap (validity, or Boolean values): tiles start on 128-row boundaries
    for (uint32_t c = 0; c < D.nstage_cols; ++c) nullable |= D.scol[c].nullable != 0 || D.scol[c].width == 0;
    auto stage_bytes_for = [&](uint32_t tile_rows) {
      uint32_t b = 0;
      for (uint32_t c = 0; c < D.nstage_cols; ++c) b += col_tile_bytes(D.scol[c], tile_rows) + (D.scol[c].nullable ? tile_rows / 8 : 0u);
      return std::max(128u, (b + 127u) & ~127u);
    };
    // Tile = the largest multiple of 32 rows (of 128 with a staged validity bitmap: its slices must start 16-byte
    // aligned) up to 256 of which two stages fit the warp's share: the per-tile work (barrier, descriptor, bulk-copy
    // issue) is amortised over more rows, and two tiles in flight per warp x 16-20 warps cover the HBM latency.
    const uint32_t step = nullable ? 128u : 32u;
    uint32_t tile_rows = 256;
    if (const char* e = std::getenv("PGF_PROBE_TILE_ROWS")) tile_rows = std::max<uint32_t>(step, uint32_t(std::atoi(e)) / step * step);   // experiments
    while (tile_rows > step && stage_bytes_for(tile_rows) * 2u > budget) tile_rows -= step;
    if (stage_bytes_for(tile_rows) * 2u > budget) {
      if (nullable) return not_eligible("row too wide for the shared-memory stages");
      tile_rows = 16;
      if (stage_bytes_for(tile_rows) * 2u > budget) return not_eligible("row too wide for the shared-memory stages");
    }
    const uint32_t stage_bytes = stage_bytes_for(tile_rows);
    uint32_t off = 0;
    for (uint32_t c = 0; c < D.nstage_cols; ++c) {
      D.scol[c].smem_off = off;
      off += col_tile_bytes(D.scol[c], tile_rows);
    }
    for (uint32_t c = 0; c < D.nstage_cols; ++c) {
      D.scol[c].valid_off = off;
      if (D.scol[c].nullable) off += tile_rows / 8;
    }
    D.nstages = std::min<uint32_t>(kPMaxDepth, budget / stage_bytes);
    if (const char* e = std::getenv("PGF_PROBE_DEPTH")) D.nstages = std::min<uint32_t>(D.nstages, std::max(2, std::atoi(e)));
    D.tile_rows = tile_rows;
    D.tiles_per_page = (max_rows + tile_rows - 1) / tile_rows;
    D.stage_bytes = stage_bytes;
    D.npages = uint32_t(s.npages);
    if (s.npages * uint64_t(D.tiles_per_page) > 0xFFFFFFF0ull) return not_eligible("scan too large for 32-bit tile ids");
    D.nitems = uint32_t(s.npages) * D.tiles_per_page;
    D.pages = s.d_pages;
    D.descs = s.d_descs;
    D.classes = s.d_classes;
    D.page_stride = ctx_->page_size;
    L_->smem = probe_shared_bytes() + size_t(kPConsumerWarps) * (size_t(D.nstages) * D.stage_bytes + kPQueueBytesPerWarp);
    return PGF_OK;
  }

  pgf_ctx* ctx_;
  const pgf_pipeline* plan_;
  bool probe_ = false;
  const JoinTable* rowset_ = nullptr;
  Lowered* L_ = nullptr;
  JoinTable* jtables_[kMaxJoins] = {nullptr, nullptr};
  int expr_cls_[kMaxExprs] = {0};
  bool expr_as_f64_[kMaxExprs] = {false};

};

// ---- result extraction ----------------------------------------------------------------
pgf_value key_value(int type, const uint64_t* words, bool is_null) {
  pgf_value v{};
  if (is_null) { v.kind = PGF_V_NULL; return v; }
  if (is_int_type(type)) {
    v.kind = PGF_V_I64;
    v.lo = int64_t(words[0]);
    v.hi = v.lo < 0 ? -1 : 0;
  } else if (type == PGF_T_DECIMAL128) {
    v.kind = PGF_V_I128;
    v.lo = int64_t(words[0]);
    v.hi = int64_t(words[1]);
  } else {  // inline view: [len u32][12 bytes]
    v.kind = PGF_V_STR;
    uint8_t raw[16];
    std::memcpy(raw, words, 16);
    uint32_t len;
    std::memcpy(&len, raw, 4);
    v.slen = int32_t(len > 12 ? 12 : len);
    std::memcpy(v.str, raw + 4, size_t(v.slen));
  }
  return v;
}

pgf_status build_result(pgf_ctx* ctx, const pgf_pipeline* plan, const Lowered& L, const std::vector<uint64_t>& state,
                        pgf_result* res) {
  const uint32_t nexprs = plan->nexprs;
  const uint32_t aw = L.acc_cls == CLS_I128 ? 2 : 1;
  const uint32_t ew = entry_words(nexprs, aw);
  const uint64_t n = state.empty() ? 0 : state[0];
  res->ngroups = n;
  res->nkeys = plan->nkeys;
  res->naggs = plan->naggs;
  for (uint32_t k = 0; k < plan->nkeys; ++k) {
    res->key_type[k] = L.key_types[k];
    res->key_not_null[k] = L.key_not_null[k] ? 1 : 0;
  }
  for (uint32_t a = 0; a < plan->naggs; ++a) {
    const pgf_agg& ag = plan->aggs[a];
    res->agg_func[a] = ag.func;
    if (ag.func == PGF_AGG_COUNT_STAR || ag.func == PGF_AGG_COUNT) res->agg_type[a] = PGF_T_INT64;
    else res->agg_type[a] = L.acc_cls == CLS_F64 ? PGF_T_FLOAT64 : L.acc_cls == CLS_I64 ? PGF_T_INT64 : PGF_T_DECIMAL128;
  }
  res->keys = new (std::nothrow) pgf_value[n * (plan->nkeys ? plan->nkeys : 1) + 1]();
  res->aggs = new (std::nothrow) pgf_value[n * (plan->naggs ? plan->naggs : 1) + 1]();
  if (!res->keys || !res->aggs) return ctx->fail(PGF_ERR_OUT_OF_MEMORY, "result allocation failed");
  for (uint64_t g = 0; g < n; ++g) {
    const uint64_t* e = state.data() + 1 + g * ew;
    const uint32_t knull = uint32_t(e[kKeyWords]);
    for (uint32_t k = 0; k < plan->nkeys; ++k)
      res->keys[g * plan->nkeys + k] = key_value(L.key_types[k], e + L.dev.keys[k].word, (knull >> k) & 1);
    const uint64_t* acc = e + kKeyWords + 1;
    const uint64_t* cnt = acc + nexprs * aw;
    for (uint32_t a = 0; a < plan->naggs; ++a) {
      pgf_value& v = res->aggs[g * plan->naggs + a];
      const pgf_agg& ag = plan->aggs[a];
      if (ag.func == PGF_AGG_COUNT_STAR) { v.kind = PGF_V_I64; v.lo = int64_t(cnt[nexprs]); continue; }
      const uint32_t x = L.expr_pos[ag.expr];   // device position of the aggregate's argument
      const uint64_t c = cnt[x];
      if (ag.func == PGF_AGG_COUNT) { v.kind = PGF_V_I64; v.lo = int64_t(c); continue; }
      if (c == 0) { v.kind = PGF_V_NULL; continue; }  // SUM / AVG over no rows is NULL
      if (L.acc_cls == CLS_F64) {
        double s;
        std::memcpy(&s, acc + x, 8);
        v.kind = PGF_V_F64;
        v.f64 = ag.func == PGF_AGG_AVG ? s / double(c) : s;  // f64 sum / (u64 count as f64)
      } else if (L.acc_cls == CLS_I64) {
        v.kind = PGF_V_I64;
        v.lo = int64_t(acc[x]);
        v.hi = v.lo < 0 ? -1 : 0;
        if (ag.func == PGF_AGG_AVG) return ctx->fail(PGF_ERR_NOT_ELIGIBLE, "integer AVG must be lowered to Float64");
      } else {
        __int128 s = (__int128)(((unsigned __int128)acc[x * 2 + 1] << 64) | acc[x * 2]);
        if (ag.func == PGF_AGG_AVG) {
          // DecimalAverager: sum * 10^(s_out - s) / count with s_out = s + 4, truncating
          s = (s * 10000) / (__int128)c;
        }
        v.kind = PGF_V_I128;
        v.lo = int64_t(uint64_t((unsigned __int128)s));
        v.hi = int64_t(uint64_t((unsigned __int128)s >> 64));
      }
    }
  }
  return PGF_OK;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

pgf_status grow(pgf_ctx* ctx, uint8_t** buf, size_t* cap, size_t need, const char* what) {
  if (*cap >= need) return PGF_OK;
  if (*buf) {
    CU(ctx, cudaStreamSynchronize(ctx->compute_stream));
    cudaFree(*buf);
    *buf = nullptr;
    *cap = 0;
  }
  const size_t want = align_up(need + need / 4, 1 << 20);
  void* p = nullptr;
  if (cudaMalloc(&p, want) != cudaSuccess) {
    cudaGetLastError();
    return ctx->fail(PGF_ERR_OUT_OF_MEMORY, "cannot allocate %zu bytes for %s", want, what);
  }
  *buf = static_cast<uint8_t*>(p);
  *cap = want;
  return PGF_OK;
}

// ---- ORDER BY / LIMIT: lowering, device top-k driver, host ordering of the returned rows ----
pgf_status lower_sort(pgf_ctx* ctx, const pgf_pipeline* plan, const Lowered& L, DevSort* S) {
  const uint32_t aw = L.acc_cls == CLS_I128 ? 2 : 1;
  S->n = plan->nsort;
  S->ew = entry_words(plan->nexprs, aw);
  if (plan->nsort > PGF_MAX_SORT) return ctx->fail(PGF_ERR_INVALID_ARGUMENT, "too many ORDER BY terms");
  const uint32_t acc0 = kKeyWords + 1, cnt0 = acc0 + plan->nexprs * aw;
  for (uint32_t i = 0; i < plan->nsort; ++i) {
    const pgf_sort_key& sk = plan->sort[i];
    DevSortKey& d = S->k[i];
    d = DevSortKey{};
    d.desc = sk.descending != 0;
    d.nulls_first = sk.nulls_first != 0;
    if (!sk.is_agg) {
      if (sk.index < 0 || uint32_t(sk.index) >= plan->nkeys) return ctx->fail(PGF_ERR_INVALID_ARGUMENT, "ORDER BY key %d out of range", sk.index);
      const int t = L.key_types[sk.index];
      d.kind = is_int_type(t) ? SK_KEY_INT : (t == PGF_T_DECIMAL128 ? SK_KEY_DEC : SK_KEY_VIEW);
      d.word = L.dev.keys[sk.index].word;
      d.null_bit = uint32_t(sk.index);
    } else {
      if (sk.index < 0 || uint32_t(sk.index) >= plan->naggs) return ctx->fail(PGF_ERR_INVALID_ARGUMENT, "ORDER BY aggregate %d out of range", sk.index);
      const pgf_agg& ag = plan->aggs[sk.index];
      if (ag.func == PGF_AGG_COUNT_STAR) { d.kind = SK_COUNT; d.cnt_word = cnt0 + plan->nexprs; continue; }
      d.word = acc0 + L.expr_pos[ag.expr] * aw;
      d.cnt_word = cnt0 + L.expr_pos[ag.expr];
      if (ag.func == PGF_AGG_COUNT) d.kind = SK_COUNT;
      else if (L.acc_cls == CLS_F64) d.kind = ag.func == PGF_AGG_AVG ? SK_F64_AVG : SK_F64_SUM;
      else if (L.acc_cls == CLS_I64) d.kind = SK_I64_SUM;
      else d.kind = ag.func == PGF_AGG_AVG ? SK_I128_AVG : SK_I128_SUM;
    }
  }
  return PGF_OK;
}

// Selects the k first entries under the ORDER BY on the device; h_state receives [k][entries].
pgf_status device_topk(pgf_ctx* ctx, const DevSort& S, const uint64_t* d_entries, uint64_t n, uint32_t k,
                       std::vector<uint64_t>* h_state, uint32_t* launches) {
  if (n >= kNoEntry) return ctx->fail(PGF_ERR_NOT_ELIGIBLE, "device top-k over %llu groups", (unsigned long long)n);
  auto blocks_for = [](uint64_t m) { return (m + kTopkSeg - 1) / kTopkSeg; };
  // candidate lists of the levels ping-pong between two buffers: level 0 writes blocks(n) * k indices, level 1 far fewer
  const uint64_t nb0 = blocks_for(n);
  const size_t o_a = 0, o_b = align_up(size_t(nb0) * k * 4, 16), o_out = o_b + align_up(size_t(blocks_for(nb0 * k)) * k * 4 + 16, 16);
  const size_t bytes = o_out + (1 + size_t(k) * S.ew) * 8;
  PGF_TRY(grow(ctx, &ctx->d_topk, &ctx->d_topk_cap, bytes, "the top-k scratch"));
  uint32_t* buf[2] = {reinterpret_cast<uint32_t*>(ctx->d_topk + o_a), reinterpret_cast<uint32_t*>(ctx->d_topk + o_b)};
  uint64_t* out = reinterpret_cast<uint64_t*>(ctx->d_topk + o_out);
  const uint32_t* in = nullptr;
  uint64_t m = n;
  int cur = 0;
  while (m > kTopkSeg) {
    const uint64_t nb = blocks_for(m);
    topk_select_kernel<<<uint32_t(nb), kTopkThreads, 0, ctx->compute_stream>>>(S, d_entries, in, uint32_t(m), k, buf[cur], nullptr);
    in = buf[cur];
    m = nb * k;
    cur ^= 1;
    ++*launches;
  }
  topk_select_kernel<<<1, kTopkThreads, 0, ctx->compute_stream>>>(S, d_entries, in, uint32_t(m), k, buf[cur], out);
  CU(ctx, cudaGetLastError());
  ++*launches;
  h_state->assign(1 + size_t(k) * S.ew, 0);
  CU(ctx, cudaMemcpyAsync(h_state->data(), out, h_state->size() * 8, cudaMemcpyDeviceToHost, ctx->compute_stream));
  CU(ctx, cudaStreamSynchronize(ctx->compute_stream));
  return PGF_OK;
}

int cmp_value(const pgf_value& a, const pgf_value& b) {  // ascending, both non-NULL
  switch (a.kind) {
    case PGF_V_F64: {
      int64_t x, y;
      std::memcpy(&x, &a.f64, 8);
      std::memcpy(&y, &b.f64, 8);
      x ^= int64_t(uint64_t(x >> 63) >> 1);  // IEEE totalOrder
      y ^= int64_t(uint64_t(y >> 63) >> 1);
      return x < y ? -1 : (x > y ? 1 : 0);
    }
    case PGF_V_I64: return a.lo < b.lo ? -1 : (a.lo > b.lo ? 1 : 0);
    case PGF_V_I128:
      if (a.hi != b.hi) return a.hi < b.hi ? -1 : 1;
      return uint64_t(a.lo) < uint64_t(b.lo) ? -1 : (uint64_t(a.lo) > uint64_t(b.lo) ? 1 : 0);
    default: {
      const int n = std::min(a.slen, b.slen);
      const int c = std::memcmp(a.str, b.str, size_t(n));
      if (c) return c < 0 ? -1 : 1;
      return a.slen < b.slen ? -1 : (a.slen > b.slen ? 1 : 0);
    }
  }
}

// Orders the rows of a result by the plan's ORDER BY and applies the LIMIT (host side: the rows
// are either few, or already the device-selected top k).
void sort_result(const pgf_pipeline* plan, pgf_result* res) {
  if (!plan->nsort && !plan->limit) return;
  const uint64_t n = res->ngroups;
  std::vector<uint64_t> idx(n);
  for (uint64_t i = 0; i < n; ++i) idx[i] = i;
  auto cell = [&](uint64_t row, const pgf_sort_key& sk) -> const pgf_value& {
    return sk.is_agg ? res->aggs[row * res->naggs + sk.index] : res->keys[row * res->nkeys + sk.index];
  };
  if (plan->nsort)
    std::stable_sort(idx.begin(), idx.end(), [&](uint64_t x, uint64_t y) {
      for (uint32_t i = 0; i < plan->nsort; ++i) {
        const pgf_sort_key& sk = plan->sort[i];
        const pgf_value &a = cell(x, sk), &b = cell(y, sk);
        const bool na = a.kind == PGF_V_NULL, nb = b.kind == PGF_V_NULL;
        if (na || nb) {
          if (na && nb) continue;
          return na == (sk.nulls_first != 0);
        }
        const int c = cmp_value(a, b);
        if (c) return sk.descending ? c > 0 : c < 0;
      }
      return false;
    });
  const uint64_t keep = plan->limit && plan->limit < n ? plan->limit : n;
  const uint32_t nk = res->nkeys ? res->nkeys : 1, na = res->naggs ? res->naggs : 1;
  pgf_value* keys = new (std::nothrow) pgf_value[keep * nk + 1]();
  pgf_value* aggs = new (std::nothrow) pgf_value[keep * na + 1]();
  if (!keys || !aggs) { delete[] keys; delete[] aggs; return; }
  for (uint64_t i = 0; i < keep; ++i) {
    for (uint32_t k = 0; k < res->nkeys; ++k) keys[i * res->nkeys + k] = res->keys[idx[i] * res->nkeys + k];
    for (uint32_t a = 0; a < res->naggs; ++a) aggs[i * res->naggs + a] = res->aggs[idx[i] * res->naggs + a];
  }
  delete[] res->keys;
  delete[] res->aggs;
  res->keys = keys;
  res->aggs = aggs;
  res->ngroups = keep;
}

// A specialised instantiation exists when every term is a plain range over the load kinds of
// a registered shape and the aggregate arguments have the registered forms.
const ShapeEntry* pick_shape(const Lowered& L) {
  const DevPlan& D = L.dev;
  if (D.nterms > 4 || D.used_null_mask != 0) return nullptr;  // registered shapes assume NOT NULL scan columns
  ShapeSig sig{};
  sig.sink = D.sink;
  sig.acc = L.acc_cls;
  sig.grouped = L.grouped;
  sig.nj = L.nj;
  sig.maxe = L.maxe;
  sig.nterms = int(D.nterms);
  for (uint32_t t = 0; t < D.nterms; ++t) {
    if (D.terms[t].op != TERM_IN_RANGE) return nullptr;
    sig.term_ld[t] = D.terms[t].ref.ld;
  }
  sig.nexprs = int(D.nexprs);
  for (uint32_t e = 0; e < D.nexprs; ++e) sig.expr_form[e] = int(D.exprs[e].form);
  sig.nkeys = int(D.nkeys);
  for (uint32_t k = 0; k < D.nkeys; ++k)
    sig.key_enc[k] = key_enc(D.keys[k].ref.ld, D.keys[k].ref.src != SRC_PAGE, D.keys[k].word);
  return find_shape(sig);
}

// Device header at the start of the arena (one memset clears header + table).
struct ArenaHeader {
  Counters counters;         // 48 bytes
  uint32_t overflow, used;   // group-table (or build-row buffer) overflow flag, occupied slots
  unsigned long long build_rows;  // rows a build sink appended (may exceed the buffer when `overflow` is set)
  uint32_t cta_done;         // CTAs that finished (fixed-order Float64 reduction of the streaming kernel)
  uint32_t entries_overflow; // split join pipeline: more tag hits than the entry buffer holds (the run is repeated fused)
  unsigned long long entries;  // ... tag hits appended
  uint32_t pad[12];
};
static_assert(sizeof(ArenaHeader) == 128, "arena header layout");
constexpr uint64_t kSmallTable = 1ull << 16;   // tables up to this many slots keep their result entries in the arena
constexpr size_t kHostArena = 64 * 1024;       // pinned mirror: header + first result entries

struct TableAlloc {
  uint64_t* d_cta_rec = nullptr;  // per-CTA Float64 sums (zeroed with the table)
  GroupTable t{};
  uint64_t capacity = 0;
  ArenaHeader* d_header = nullptr;
  size_t zero_bytes = 0;     // header + table
  uint64_t* d_out = nullptr; // [count][entries] (small tables only)
  uint64_t out_entries = 0;
};

// Places header, group table and (for small tables) the result entries in the context's
// grow-only arena and clears header + table with one memset on the compute stream.
pgf_status arena_table(pgf_ctx* ctx, uint64_t capacity, uint32_t nexprs, uint32_t acc_words, TableAlloc* out) {
  const uint32_t ne = nexprs ? nexprs : 1;
  const uint32_t ew = entry_words(nexprs, acc_words);
  // [header | per-CTA records | state] are cleared; keys / accumulators / counts of a large table are not: the
  // thread that creates a group initialises its slot (group_slot_init).  Small tables are cleared whole -- it is
  // free, and the single slot of an aggregate without GROUP BY is never "created".
  size_t off = sizeof(ArenaHeader);
  const size_t o_rec = off;   off += size_t(ctx->sm_count) * kRegGroups * (2 + nexprs) * 8;
  const size_t o_state = off = align_up(off, 16); off = align_up(off + capacity * 4, 16);   // (16-byte aligned: the extract kernel reads four state words per load)
  const size_t zero_large = off;
  const size_t o_keys = off;  off += capacity * kKeyWords * 8;
  const size_t o_acc = off;   off += capacity * ne * acc_words * 8;
  const size_t o_cnt = off;   off += capacity * (nexprs + 1) * 8;
  out->zero_bytes = capacity > (1u << 16) ? zero_large : off;
  const bool small = capacity <= kSmallTable;
  const size_t o_out = off = align_up(off, 16);
  if (small) off += (1 + capacity * ew) * 8;
  PGF_TRY(grow(ctx, &ctx->d_arena, &ctx->d_arena_cap, off, "the group table"));
  uint8_t* base = ctx->d_arena;
  out->capacity = capacity;
  out->d_header = reinterpret_cast<ArenaHeader*>(base);
  GroupTable& t = out->t;
  t.state = reinterpret_cast<uint32_t*>(base + o_state);
  t.keys = reinterpret_cast<uint64_t*>(base + o_keys);
  t.acc = reinterpret_cast<uint64_t*>(base + o_acc);
  t.cnt = reinterpret_cast<uint64_t*>(base + o_cnt);
  t.mask = uint32_t(capacity - 1);
  t.acc_words = acc_words;
  t.nexprs = nexprs;
  t.overflow = &out->d_header->overflow;
  t.used = &out->d_header->used;
  out->d_cta_rec = reinterpret_cast<uint64_t*>(base + o_rec);
  out->d_out = small ? reinterpret_cast<uint64_t*>(base + o_out) : nullptr;
  out->out_entries = small ? capacity : 0;
  CU(ctx, cudaMemsetAsync(base, 0, out->zero_bytes, ctx->compute_stream));
  if (small) CU(ctx, cudaMemsetAsync(out->d_out, 0, 8, ctx->compute_stream));
  return PGF_OK;
}

// Extract the occupied slots into `d_state` ([count][entries...]) on the compute stream;
// d_state[0] must already be zero.
pgf_status extract_table(pgf_ctx* ctx, const GroupTable& t, uint64_t capacity, uint32_t nexprs, bool grouped,
                         uint64_t* d_state, uint64_t max_entries) {
  const uint64_t nslots = grouped ? capacity : 1;
  const uint32_t grid = uint32_t(std::min<uint64_t>((nslots + kExtractThreads - 1) / kExtractThreads, uint64_t(ctx->sm_count) * 8));
  table_extract_kernel<<<grid, kExtractThreads, 0, ctx->compute_stream>>>(t, nexprs, grouped, d_state, max_entries);
  CU(ctx, cudaGetLastError());
  return PGF_OK;
}

// Launch the fused kernel the lowering chose: the compaction pipeline, a registered shape, or the generic
// streaming instantiation.
cudaError_t launch_fused(const Lowered& L, uint32_t grid, cudaStream_t stream) {
  if (L.rowscan) return launch_rows(L.acc_cls, L.dev, grid, stream);
  if (L.probe && L.split) return launch_probe_split(L.acc_cls, L.t0, L.dev, grid, L.smem, L.split_grid, stream);
  if (L.probe) return launch_probe(L.acc_cls, L.t0, L.dev, grid, L.smem, stream);
  if (const ShapeEntry* se = pick_shape(L)) return se->fn(L.dev, grid, L.smem, stream);
  return launch_pipeline(L.dev.sink, L.acc_cls, L.grouped, L.nj, L.maxe, L.dev, grid, L.smem, stream);
}
const char* variant_name(const Lowered& L) {
  if (L.rowscan) return "row_set_scan";
  if (L.probe) return L.t0 == LD_VIEW ? "compact_1_string_term" : "compact_generic";
  const ShapeEntry* se = pick_shape(L);
  return se ? se->name : "generic";
}

}  // namespace

// RuntimeFilter* counters of a fused run (runtime_metrics/src/lib.rs:128-131; caller holds ctx->mu): rows tested
// against a filter -- every scanned row for filters probed in stage A, the rows past the predicate for the dense
// filter of the compaction pipeline --, rows rejected, and rows that met a probe the lowering had dropped.
static void note_fused_probes(pgf_ctx* ctx, const Lowered& L, const Counters& c, uint64_t build_rows) {
  pgf_runtime_filter_metrics& m = ctx->rf_metrics;
  m.build_rows_total += build_rows;
  const uint64_t rejected = c.rows_in - c.rows_bloom;
  if (L.dev.nbloom) {
    const bool dense_only = L.dev.bloom_dense && L.dev.nbloom == 1;
    m.probe_rows_total += dense_only ? c.rows_filtered + rejected : c.rows_in;
    m.probe_rows_rejected_total += rejected;
  }
  m.probe_pass_unfiltered_total += uint64_t(L.bloom_dropped) * c.rows_in;
}

pgf_status pipeline_run(pgf_ctx* ctx, const pgf_pipeline* plan, bool check_only, void* dev_state_out,
                        uint64_t state_cap, uint64_t* state_bytes, bool partial, pgf_result** out) {
  NvtxRange nvtx_("pgf:pipeline_run");
  std::lock_guard<std::mutex> g(ctx->mu);
  CU(ctx, cudaSetDevice(ctx->device));
  Lowered L;
  Lowering low(ctx, plan);
  PGF_TRY(low.run(&L));
  if (plan->sink == PGF_SINK_AGGREGATE && plan->nsort) {  // ORDER BY terms are part of the eligibility check
    DevSort S;
    PGF_TRY(lower_sort(ctx, plan, L, &S));
  }
  if (check_only) return PGF_OK;
  if (partial && plan->sink != PGF_SINK_AGGREGATE) return ctx->fail(PGF_ERR_INVALID_ARGUMENT, "partial states exist for aggregate sinks only");
  if (!L.rowscan) {
    PGF_TRY(scan_sync_descs(ctx, *L.scan));
    L.dev.descs = L.scan->d_descs;
    L.dev.classes = L.scan->d_classes;
    if (L.scan->h_classes.size() == 1) {
      L.dev.single_class = 1;
      L.dev.class0 = L.scan->h_classes[0];
    }
  }

  pgf_result* res = new (std::nothrow) pgf_result();
  if (!res) return ctx->fail(PGF_ERR_OUT_OF_MEMORY, "result allocation failed");
  struct ResGuard {
    pgf_result* r;
    ~ResGuard() { if (r) pgf_result_free(r); }
  } guard{res};

  DevAlloc mem;
  const bool agg = plan->sink == PGF_SINK_AGGREGATE;
  const uint32_t aw = L.acc_cls == CLS_I128 ? 2 : 1;
  const uint32_t ew = entry_words(plan->nexprs, aw);
  uint32_t grid = uint32_t(std::min<uint64_t>(L.dev.npages ? L.dev.npages : 1, uint64_t(ctx->sm_count)));  // CTAs take whole pages
  if (L.probe)  // the compaction pipeline deals tiles to warps: one CTA per SM while every warp has a tile
    grid = uint32_t(std::min<uint64_t>(std::max<uint64_t>((uint64_t(L.dev.nitems) + kPConsumerWarps - 1) / kPConsumerWarps, 1), uint64_t(ctx->sm_count)));
  if (L.rowscan) grid = uint32_t(std::min<uint64_t>(std::max<uint64_t>((L.dev.row_count + 255) / 256, 1), uint64_t(ctx->sm_count) * 8));
  uint64_t capacity = L.table_capacity;
  float total_ms = 0.f;
  uint32_t launches = 0;
  ArenaHeader* h_header = reinterpret_cast<ArenaHeader*>(ctx->h_arena);
  uint64_t* h_out = reinterpret_cast<uint64_t*>(ctx->h_arena + sizeof(ArenaHeader));
  const uint64_t h_out_entries = (kHostArena / 2 - sizeof(ArenaHeader) - 8) / (uint64_t(ew) * 8);

  // HashJoinExec build side: the fused kernel appends the build rows to a dense buffer; the table is
  // built from them afterwards, sized by the rows that actually arrived (not by the rows scanned), so
  // its tag directory is as small -- as L2 resident -- as the join allows.
  struct RowBuf {
    pgf_ctx* ctx;
    void* p = nullptr;
    size_t bytes = 0;
    ~RowBuf() { if (p) ctx->join_free(p, bytes); }
  } rowbuf{ctx};
  uint64_t rows_cap = 0, bloom_rows = 0;
  if (plan->sink == PGF_SINK_JOIN_BUILD) {
    rows_cap = plan->expected_groups ? plan->expected_groups : std::max<uint64_t>(L.scan->rows, 1);
    rowbuf.p = ctx->join_alloc(rows_cap * L.build_table.slot_u4 * sizeof(uint4), &rowbuf.bytes);
    if (!rowbuf.p) return ctx->fail(PGF_ERR_OUT_OF_MEMORY, "cannot allocate a build-row buffer of %llu rows", (unsigned long long)rows_cap);
  }

  // Aggregates behind one join probe run split (stages A + B, then stage C over the tag hits) unless
  // PGF_PROBE_SPLIT=0.  The entry buffer has room for every scanned row (16 bytes each, grow-only, shared by the
  // context's plans): a row appends at most one entry.  Should the buffer be unobtainable the plan runs fused; a
  // smaller buffer that overflows (tests force one) repeats the run fused.  (Build sinks stay fused: nearly all their
  // entries are true matches that end in one append cursor, and the split was 5 % slower there.)
  uint64_t entries_cap = 0;
  if (L.probe && !L.rowscan && L.dev.njoins == 1 && L.dev.sink == SINK_AGG && L.dev.nitems) {
    const char* e = std::getenv("PGF_PROBE_SPLIT");
    if (!e || *e != '0') {
      entries_cap = std::max<uint64_t>(L.scan->rows, 1u << 10);
      if (const char* c = std::getenv("PGF_PROBE_SPLIT_CAP")) entries_cap = std::max<uint64_t>(32, uint64_t(std::atoll(c)));   // tests: force the fallback
      if (grow(ctx, &ctx->d_entries, &ctx->d_entries_cap, entries_cap * sizeof(uint4), "the stage-C entry buffer") == PGF_OK) {
        L.split = true;
        L.split_grid = uint32_t(ctx->sm_count) * 4u;
      }   // (else: out of memory for an optional buffer is not an error of the plan -- it runs fused)
    }
  }

  PhaseTrace trace(ctx->compute_stream);
  for (int attempt = 0; attempt < 6; ++attempt) {
    TableAlloc ta;
    // header (+ group table + result entries) live in the arena; other sinks only use the header
    PGF_TRY(arena_table(ctx, agg ? capacity : 1, agg ? plan->nexprs : 0, aw, &ta));
    trace.mark("arena + clear");
    L.dev.table = ta.t;
    L.dev.counters = &ta.d_header->counters;
    L.dev.build_count = &ta.d_header->build_rows;
    L.dev.cta_rec = ta.d_cta_rec;
    L.dev.cta_done = &ta.d_header->cta_done;
    L.dev.build.rows = static_cast<uint4*>(rowbuf.p);
    L.dev.build.rows_cap = rows_cap;
    L.dev.entries = reinterpret_cast<uint4*>(ctx->d_entries);
    L.dev.entries_cap = entries_cap;
    L.dev.entries_count = &ta.d_header->entries;
    L.dev.entries_overflow = &ta.d_header->entries_overflow;
    CU(ctx, cudaEventRecord(ctx->ev_a, ctx->compute_stream));
    if (L.dev.nitems) {
      CU(ctx, launch_fused(L, grid, ctx->compute_stream));
      launches += L.split ? 2 : 1;
    }
    CU(ctx, cudaEventRecord(ctx->ev_b, ctx->compute_stream));
    trace.mark("fused kernel");
    // small tables: extract right away so header and result travel with one synchronisation
    const bool inline_out = agg && !partial && ta.d_out != nullptr;
    uint64_t prefix_entries = 0;
    if (inline_out) {
      PGF_TRY(extract_table(ctx, ta.t, capacity, plan->nexprs, L.grouped, ta.d_out, ta.out_entries));
      ++launches;
      prefix_entries = std::min<uint64_t>(ta.out_entries, h_out_entries);
      CU(ctx, cudaMemcpyAsync(h_out, ta.d_out, (1 + prefix_entries * ew) * 8, cudaMemcpyDeviceToHost, ctx->compute_stream));
    }
    CU(ctx, cudaMemcpyAsync(h_header, ta.d_header, sizeof(ArenaHeader), cudaMemcpyDeviceToHost, ctx->compute_stream));
    CU(ctx, cudaStreamSynchronize(ctx->compute_stream));
    float ms = 0.f;
    CU(ctx, cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b));
    total_ms += ms;
    if (h_header->counters.bad_rows)
      return ctx->fail(PGF_ERR_UNSUPPORTED_DATA, "%llu rows carry out-of-line (> 12 byte) view values in a predicate or key column",
                       (unsigned long long)h_header->counters.bad_rows);
    if (L.split && h_header->entries_overflow) {  // more tag hits than the entry buffer holds: one fused kernel instead
      L.split = false;
      continue;
    }
    if (agg && h_header->overflow) {  // group table overflow: grow and re-run
      capacity *= 16;
      if (capacity > (1ull << 30)) return ctx->fail(PGF_ERR_OUT_OF_MEMORY, "group table would exceed 2^30 slots");
      continue;
    }
    if (plan->sink == PGF_SINK_JOIN_BUILD) {
      const uint64_t nrows = h_header->build_rows;
      if (h_header->overflow) {  // more build rows than the buffer holds (a hint that was too small, or a join that multiplies rows)
        ctx->join_free(rowbuf.p, rowbuf.bytes);
        rowbuf.p = nullptr;
        rows_cap = nrows + nrows / 8 + 1024;  // the kernel counted every row it wanted to append
        rowbuf.p = ctx->join_alloc(rows_cap * L.build_table.slot_u4 * sizeof(uint4), &rowbuf.bytes);
        if (!rowbuf.p) return ctx->fail(PGF_ERR_OUT_OF_MEMORY, "cannot allocate a build-row buffer of %llu rows", (unsigned long long)rows_cap);
        continue;
      }
      JoinTable& jt = L.build_table;
      jt.rows = nrows;
      if (L.dev.has_build_bloom && nrows) {
        CU(ctx, cudaEventRecord(ctx->ev_a, ctx->compute_stream));
        const uint32_t bgrid = uint32_t(std::min<uint64_t>((nrows + 255) / 256, uint64_t(ctx->sm_count) * 8));
        bloom_insert_rows_kernel<<<bgrid, 256, 0, ctx->compute_stream>>>(L.dev.build_bloom, static_cast<const uint4*>(rowbuf.p), nrows, jt.slot_u4);
        CU(ctx, cudaGetLastError());
        CU(ctx, cudaEventRecord(ctx->ev_b, ctx->compute_stream));
        CU(ctx, cudaEventSynchronize(ctx->ev_b));
        CU(ctx, cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b));
        total_ms += ms;
        ++launches;
        bloom_rows = nrows;
      }
      if (plan->build_flags & PGF_BUILD_ROWS_ONLY) {  // the row set is the result: no hash table
        jt.d_rows = static_cast<uint4*>(rowbuf.p);
        jt.rows_alloc_bytes = rowbuf.bytes;
        rowbuf.p = nullptr;
        jt.capacity = 0;
        break;
      }
      uint64_t cap = 1024;
      while (cap < nrows * 2) cap <<= 1;
      if (cap > (1ull << 31)) return ctx->fail(PGF_ERR_NOT_ELIGIBLE, "join build side too large for one table");
      jt.capacity = uint32_t(cap);
      jt.rows = nrows;
      // slots followed by the one-byte tag directory
      const uint64_t slot_bytes = cap * jt.slot_u4 * sizeof(uint4), tag_bytes = cap + 16;
      jt.d_slots = static_cast<uint4*>(ctx->join_alloc(slot_bytes + tag_bytes, &jt.alloc_bytes));
      if (!jt.d_slots) return ctx->fail(PGF_ERR_OUT_OF_MEMORY, "cannot allocate a join table of %u slots", jt.capacity);
      mem.ptrs.push_back(jt.d_slots);
      CU(ctx, cudaEventRecord(ctx->ev_a, ctx->compute_stream));
      CU(ctx, cudaMemsetAsync(reinterpret_cast<uint8_t*>(jt.d_slots) + slot_bytes, 0, tag_bytes, ctx->compute_stream));
      if (nrows) {
        uint32_t shift = 64;
        for (uint64_t c = cap; c > 1; c >>= 1) shift--;
        const uint32_t bgrid = uint32_t(std::min<uint64_t>((nrows + 255) / 256, uint64_t(ctx->sm_count) * 8));
        join_build_kernel<<<bgrid, 256, 0, ctx->compute_stream>>>(jt.d_slots, reinterpret_cast<uint8_t*>(jt.d_slots) + slot_bytes, jt.capacity - 1,
                                                                  shift, jt.slot_u4, static_cast<const uint4*>(rowbuf.p), nrows);
        CU(ctx, cudaGetLastError());
        ++launches;
      }
      CU(ctx, cudaEventRecord(ctx->ev_b, ctx->compute_stream));
      CU(ctx, cudaEventSynchronize(ctx->ev_b));
      CU(ctx, cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b));
      total_ms += ms;  // clearing + filling the table is part of the build side's device time
      trace.mark("join table from rows");
    }
    if (agg) {
      const uint64_t ngroups = L.grouped ? h_header->used : 1;
      const uint64_t bytes = (1 + ngroups * ew) * 8;
      if (partial) {
        if (bytes > state_cap) return ctx->fail(PGF_ERR_INVALID_ARGUMENT, "partial state needs %llu bytes, buffer has %llu", (unsigned long long)bytes, (unsigned long long)state_cap);
        CU(ctx, cudaMemsetAsync(dev_state_out, 0, 8, ctx->compute_stream));
        PGF_TRY(extract_table(ctx, ta.t, capacity, plan->nexprs, L.grouped, static_cast<uint64_t*>(dev_state_out), ngroups));
        CU(ctx, cudaStreamSynchronize(ctx->compute_stream));
        *state_bytes = bytes;
        ++launches;
      } else {
        const bool topk = plan->nsort && plan->limit && plan->limit <= PGF_TOPK_DEVICE_MAX && ngroups > plan->limit;
        std::vector<uint64_t> h_state;
        if (!topk) h_state.resize(1 + ngroups * ew);  // (an 80 MB host buffer at SF100: only when every group is returned)
        if (topk) {
          // TopK: only `limit` rows leave the GPU
          const uint64_t* d_entries = nullptr;
          if (inline_out) {
            d_entries = ta.d_out + 1;
          } else {
            PGF_TRY(grow(ctx, &ctx->d_out, &ctx->d_out_cap, bytes, "the result buffer"));
            uint64_t* d_state = reinterpret_cast<uint64_t*>(ctx->d_out);
            CU(ctx, cudaMemsetAsync(d_state, 0, 8, ctx->compute_stream));
            PGF_TRY(extract_table(ctx, ta.t, capacity, plan->nexprs, L.grouped, d_state, ngroups));
            ++launches;
            d_entries = d_state + 1;
          }
          trace.mark("extract groups");
          DevSort S;
          PGF_TRY(lower_sort(ctx, plan, L, &S));
          PGF_TRY(device_topk(ctx, S, d_entries, ngroups, uint32_t(plan->limit), &h_state, &launches));
          trace.mark("device top-k");
          if (h_state[0] != plan->limit) return ctx->fail(PGF_ERR_STATE, "top-k selection returned %llu rows", (unsigned long long)h_state[0]);
        } else if (inline_out) {
          const uint64_t have = std::min<uint64_t>(ngroups, prefix_entries);
          std::memcpy(h_state.data(), h_out, (1 + have * ew) * 8);
          if (ngroups > have) {
            CU(ctx, cudaMemcpyAsync(h_state.data() + 1 + have * ew, ta.d_out + 1 + have * ew, (ngroups - have) * ew * 8,
                                    cudaMemcpyDeviceToHost, ctx->compute_stream));
            CU(ctx, cudaStreamSynchronize(ctx->compute_stream));
          }
        } else {
          PGF_TRY(grow(ctx, &ctx->d_out, &ctx->d_out_cap, bytes, "the result buffer"));
          uint64_t* d_state = reinterpret_cast<uint64_t*>(ctx->d_out);
          CU(ctx, cudaMemsetAsync(d_state, 0, 8, ctx->compute_stream));
          PGF_TRY(extract_table(ctx, ta.t, capacity, plan->nexprs, L.grouped, d_state, ngroups));
          CU(ctx, cudaMemcpyAsync(h_state.data(), d_state, bytes, cudaMemcpyDeviceToHost, ctx->compute_stream));
          CU(ctx, cudaStreamSynchronize(ctx->compute_stream));
          ++launches;
        }
        if (!topk && h_state[0] != ngroups) return ctx->fail(PGF_ERR_STATE, "group table extraction found %llu groups, expected %llu",
                                                             (unsigned long long)h_state[0], (unsigned long long)ngroups);
        PGF_TRY(build_result(ctx, plan, L, h_state, res));
        sort_result(plan, res);
      }
    }
    break;
  }

  trace.mark("result to host");
  const Counters& c = h_header->counters;
  res->rows_in = c.rows_in;
  res->rows_bloom = c.rows_bloom;
  res->rows_filtered = c.rows_filtered;
  res->rows_out = c.rows_out;
  res->bloom_rows = bloom_rows;
  note_fused_probes(ctx, L, c, bloom_rows);
  res->kernel_ms = total_ms;
  ctx->last_kernel_ms = total_ms;
  res->kernel_launches = launches;
  std::snprintf(res->variant, sizeof res->variant, "%s", variant_name(L));
  if (plan->sink == PGF_SINK_JOIN_BUILD) {
    JoinTable jt = L.build_table;
    if (jt.d_slots) mem.release(jt.d_slots);
    const uint64_t id = ctx->next_handle++;
    ctx->joins[id] = jt;
    res->join_table = id;
  }
  if (out) {
    *out = res;
    guard.r = nullptr;
  }
  return PGF_OK;
}

// Asynchronous partial run for the multi-GPU step: fused kernel + extraction of the partial
// state are enqueued on the compute stream; nothing is synchronised.  The run's header
// (counters, overflow flag) is copied to the second half of the pinned mirror and checked by
// the bounded merge that follows on the same stream.
constexpr size_t kPartialHeaderOff = kHostArena / 2;

pgf_status pipeline_run_partial_async(pgf_ctx* ctx, const pgf_pipeline* plan, void* dev_state_out, uint64_t state_cap) {
  NvtxRange nvtx_("pgf:pipeline_partial");
  std::lock_guard<std::mutex> g(ctx->mu);
  CU(ctx, cudaSetDevice(ctx->device));
  if (plan->sink != PGF_SINK_AGGREGATE) return ctx->fail(PGF_ERR_INVALID_ARGUMENT, "partial states exist for aggregate sinks only");
  Lowered L;
  Lowering low(ctx, plan);
  PGF_TRY(low.run(&L));
  if (!L.rowscan) {
    PGF_TRY(scan_sync_descs(ctx, *L.scan));
    L.dev.descs = L.scan->d_descs;
    L.dev.classes = L.scan->d_classes;
    if (L.scan->h_classes.size() == 1) {
      L.dev.single_class = 1;
      L.dev.class0 = L.scan->h_classes[0];
    }
  }
  const uint32_t aw = L.acc_cls == CLS_I128 ? 2 : 1;
  const uint32_t ew = entry_words(plan->nexprs, aw);
  if (state_cap < (1 + uint64_t(ew)) * 8) return ctx->fail(PGF_ERR_INVALID_ARGUMENT, "partial state buffer too small");
  const uint64_t max_entries = (state_cap / 8 - 1) / ew;
  TableAlloc ta;
  PGF_TRY(arena_table(ctx, L.table_capacity, plan->nexprs, aw, &ta));
  L.dev.table = ta.t;
  L.dev.counters = &ta.d_header->counters;
  uint32_t grid = uint32_t(std::min<uint64_t>(L.dev.npages ? L.dev.npages : 1, uint64_t(ctx->sm_count)));  // CTAs take whole pages
  if (L.probe)
    grid = uint32_t(std::min<uint64_t>(std::max<uint64_t>((uint64_t(L.dev.nitems) + kPConsumerWarps - 1) / kPConsumerWarps, 1), uint64_t(ctx->sm_count)));
  if (L.rowscan) grid = uint32_t(std::min<uint64_t>(std::max<uint64_t>((L.dev.row_count + 255) / 256, 1), uint64_t(ctx->sm_count) * 8));
  L.dev.build_count = &ta.d_header->build_rows;
  L.dev.cta_rec = ta.d_cta_rec;
  L.dev.cta_done = &ta.d_header->cta_done;
  CU(ctx, cudaEventRecord(ctx->ev_a, ctx->compute_stream));
  if (L.dev.nitems) CU(ctx, launch_fused(L, grid, ctx->compute_stream));
  CU(ctx, cudaEventRecord(ctx->ev_b, ctx->compute_stream));
  CU(ctx, cudaMemsetAsync(dev_state_out, 0, 8, ctx->compute_stream));
  PGF_TRY(extract_table(ctx, ta.t, L.table_capacity, plan->nexprs, L.grouped, static_cast<uint64_t*>(dev_state_out), max_entries));
  CU(ctx, cudaMemcpyAsync(ctx->h_arena + kPartialHeaderOff, ta.d_header, sizeof(ArenaHeader), cudaMemcpyDeviceToHost, ctx->compute_stream));
  ctx->partial_pending = true;
  return PGF_OK;
}

pgf_status pipeline_merge(pgf_ctx* ctx, const pgf_pipeline* plan, const void* dev_states, uint64_t stride,
                          uint32_t nstates, bool bounded, pgf_result** out) {
  NvtxRange nvtx_("pgf:merge_partials");
  std::lock_guard<std::mutex> g(ctx->mu);
  CU(ctx, cudaSetDevice(ctx->device));
  if (plan->sink != PGF_SINK_AGGREGATE) return ctx->fail(PGF_ERR_INVALID_ARGUMENT, "partial states exist for aggregate sinks only");
  Lowered L;
  Lowering low(ctx, plan);
  PGF_TRY(low.run(&L));
  if (plan->nsort) {  // validate the ORDER BY terms before anything is ordered by them
    DevSort S;
    PGF_TRY(lower_sort(ctx, plan, L, &S));
  }
  const uint32_t aw = L.acc_cls == CLS_I128 ? 2 : 1;
  const uint32_t ew = entry_words(plan->nexprs, aw);
  if (stride < (1 + uint64_t(ew)) * 8) return ctx->fail(PGF_ERR_INVALID_ARGUMENT, "partial state stride too small");
  const uint64_t max_entries = (stride / 8 - 1) / ew;
  // size the final table from the partial group counts (bounded: from what the strides can hold,
  // without reading the counts back -- the whole multi-GPU step then synchronises once)
  std::vector<uint64_t> counts(nstates, max_entries);
  uint64_t total = 0;
  if (!bounded) {
    for (uint32_t i = 0; i < nstates; ++i) {
      CU(ctx, cudaMemcpyAsync(&counts[i], static_cast<const uint8_t*>(dev_states) + i * stride, 8, cudaMemcpyDeviceToHost, ctx->compute_stream));
    }
    CU(ctx, cudaStreamSynchronize(ctx->compute_stream));
  }
  for (uint32_t i = 0; i < nstates; ++i) {
    if (counts[i] > max_entries) return ctx->fail(PGF_ERR_INVALID_ARGUMENT, "partial state %u overruns its stride", i);
    total += counts[i];
  }
  uint64_t capacity = 1;
  if (L.grouped) {
    capacity = 1024;
    while (capacity < total * 2) capacity <<= 1;
  }
  TableAlloc ta;
  PGF_TRY(arena_table(ctx, capacity, plan->nexprs, aw, &ta));
  const bool one_launch = bounded && max_entries <= 1024;
  if (one_launch) {
    const uint8_t* st = static_cast<const uint8_t*>(dev_states);
    switch (L.acc_cls) {
      case CLS_F64: table_merge_all_kernel<CLS_F64><<<1, 256, 0, ctx->compute_stream>>>(ta.t, plan->nexprs, L.dev.nkeywords, st, stride, nstates, max_entries); break;
      case CLS_I64: table_merge_all_kernel<CLS_I64><<<1, 256, 0, ctx->compute_stream>>>(ta.t, plan->nexprs, L.dev.nkeywords, st, stride, nstates, max_entries); break;
      default: table_merge_all_kernel<CLS_I128><<<1, 256, 0, ctx->compute_stream>>>(ta.t, plan->nexprs, L.dev.nkeywords, st, stride, nstates, max_entries); break;
    }
    CU(ctx, cudaGetLastError());
  }
  for (uint32_t i = 0; i < nstates && !one_launch; ++i) {  // rank order => fixed Float64 summation order
    if (!counts[i]) continue;
    const uint64_t* st = reinterpret_cast<const uint64_t*>(static_cast<const uint8_t*>(dev_states) + i * stride);
    const uint32_t grid = uint32_t(std::min<uint64_t>((counts[i] + 255) / 256, uint64_t(ctx->sm_count) * 4));
    switch (L.acc_cls) {
      case CLS_F64: table_merge_kernel<CLS_F64><<<grid, 256, 0, ctx->compute_stream>>>(ta.t, plan->nexprs, L.dev.nkeywords, st, max_entries); break;
      case CLS_I64: table_merge_kernel<CLS_I64><<<grid, 256, 0, ctx->compute_stream>>>(ta.t, plan->nexprs, L.dev.nkeywords, st, max_entries); break;
      default: table_merge_kernel<CLS_I128><<<grid, 256, 0, ctx->compute_stream>>>(ta.t, plan->nexprs, L.dev.nkeywords, st, max_entries); break;
    }
    CU(ctx, cudaGetLastError());
  }
  ArenaHeader* h_header = reinterpret_cast<ArenaHeader*>(ctx->h_arena);
  uint64_t* h_out = reinterpret_cast<uint64_t*>(ctx->h_arena + sizeof(ArenaHeader));
  const uint64_t h_out_entries = (kPartialHeaderOff - sizeof(ArenaHeader) - 8) / (uint64_t(ew) * 8);
  uint64_t prefix_entries = 0;
  if (ta.d_out) {  // small table: extract now, header and result travel with one synchronisation
    PGF_TRY(extract_table(ctx, ta.t, capacity, plan->nexprs, L.grouped, ta.d_out, ta.out_entries));
    prefix_entries = std::min<uint64_t>(ta.out_entries, h_out_entries);
    CU(ctx, cudaMemcpyAsync(h_out, ta.d_out, (1 + prefix_entries * ew) * 8, cudaMemcpyDeviceToHost, ctx->compute_stream));
  }
  CU(ctx, cudaMemcpyAsync(h_header, ta.d_header, sizeof(ArenaHeader), cudaMemcpyDeviceToHost, ctx->compute_stream));
  CU(ctx, cudaStreamSynchronize(ctx->compute_stream));
  const ArenaHeader* ph = reinterpret_cast<const ArenaHeader*>(ctx->h_arena + kPartialHeaderOff);
  const bool had_partial = ctx->partial_pending;
  ctx->partial_pending = false;
  if (had_partial) {
    if (ph->counters.bad_rows)
      return ctx->fail(PGF_ERR_UNSUPPORTED_DATA, "%llu rows carry out-of-line (> 12 byte) view values in a predicate or key column",
                       (unsigned long long)ph->counters.bad_rows);
    if (ph->overflow) return ctx->fail(PGF_ERR_STATE, "group table overflow in the partial run: pass a larger expected_groups");
  }
  if (h_header->overflow == 2u) return ctx->fail(PGF_ERR_INVALID_ARGUMENT, "a partial state holds more groups than its stride can carry");
  if (h_header->overflow) return ctx->fail(PGF_ERR_OUT_OF_MEMORY, "final group table overflow");
  const uint64_t ngroups = L.grouped ? h_header->used : 1;
  const uint64_t bytes = (1 + ngroups * ew) * 8;
  const bool topk = plan->nsort && plan->limit && plan->limit <= PGF_TOPK_DEVICE_MAX && ngroups > plan->limit;
  std::vector<uint64_t> h_state;
  if (!topk) h_state.resize(1 + ngroups * ew);
  uint32_t topk_launches = 0;
  if (topk) {  // ORDER BY ... LIMIT k of the merged groups: select on the device, k rows leave the GPU
    const uint64_t* d_entries = nullptr;
    if (ta.d_out) {
      d_entries = ta.d_out + 1;
    } else {
      PGF_TRY(grow(ctx, &ctx->d_out, &ctx->d_out_cap, bytes, "the result buffer"));
      uint64_t* d_state = reinterpret_cast<uint64_t*>(ctx->d_out);
      CU(ctx, cudaMemsetAsync(d_state, 0, 8, ctx->compute_stream));
      PGF_TRY(extract_table(ctx, ta.t, capacity, plan->nexprs, L.grouped, d_state, ngroups));
      d_entries = d_state + 1;
    }
    DevSort S;
    PGF_TRY(lower_sort(ctx, plan, L, &S));
    PGF_TRY(device_topk(ctx, S, d_entries, ngroups, uint32_t(plan->limit), &h_state, &topk_launches));
  } else if (ta.d_out) {
    const uint64_t have = std::min<uint64_t>(ngroups, prefix_entries);
    std::memcpy(h_state.data(), h_out, (1 + have * ew) * 8);
    if (ngroups > have) {
      CU(ctx, cudaMemcpyAsync(h_state.data() + 1 + have * ew, ta.d_out + 1 + have * ew, (ngroups - have) * ew * 8,
                              cudaMemcpyDeviceToHost, ctx->compute_stream));
      CU(ctx, cudaStreamSynchronize(ctx->compute_stream));
    }
  } else {
    PGF_TRY(grow(ctx, &ctx->d_out, &ctx->d_out_cap, bytes, "the result buffer"));
    uint64_t* d_state = reinterpret_cast<uint64_t*>(ctx->d_out);
    CU(ctx, cudaMemsetAsync(d_state, 0, 8, ctx->compute_stream));
    PGF_TRY(extract_table(ctx, ta.t, capacity, plan->nexprs, L.grouped, d_state, ngroups));
    CU(ctx, cudaMemcpyAsync(h_state.data(), d_state, bytes, cudaMemcpyDeviceToHost, ctx->compute_stream));
    CU(ctx, cudaStreamSynchronize(ctx->compute_stream));
  }
  pgf_result* res = new (std::nothrow) pgf_result();
  if (!res) return ctx->fail(PGF_ERR_OUT_OF_MEMORY, "result allocation failed");
  pgf_status st = build_result(ctx, plan, L, h_state, res);
  if (st) {
    pgf_result_free(res);
    return st;
  }
  sort_result(plan, res);  // ORDER BY / LIMIT of the final (merged) result
  res->kernel_launches = nstates + 1 + topk_launches;
  if (had_partial) {  // statistics of the asynchronous partial run that fed this merge
    const Counters& c = ph->counters;
    note_fused_probes(ctx, L, c, 0);
    res->rows_in = c.rows_in;
    res->rows_bloom = c.rows_bloom;
    res->rows_filtered = c.rows_filtered;
    res->rows_out = c.rows_out;
    float ms = 0.f;
    CU(ctx, cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b));
    res->kernel_ms = ms;
    ctx->last_kernel_ms = ms;
    res->kernel_launches += 2;
  }
  *out = res;
  return PGF_OK;
}

pgf_status join_export(pgf_ctx* ctx, const JoinTable& jt, void* dev_rows_out, uint64_t capacity_rows, uint64_t* rows_out) {
  std::lock_guard<std::mutex> g(ctx->mu);
  CU(ctx, cudaSetDevice(ctx->device));
  unsigned long long* d_cnt = reinterpret_cast<unsigned long long*>(ctx->d_flags + 4);
  unsigned long long* h_cnt = reinterpret_cast<unsigned long long*>(ctx->h_flags + 4);
  CU(ctx, cudaMemsetAsync(d_cnt, 0, 8, ctx->compute_stream));
  const uint32_t grid = uint32_t(std::min<uint64_t>((uint64_t(jt.capacity) + kExtractThreads - 1) / kExtractThreads, uint64_t(ctx->sm_count) * 8));
  join_export_kernel<<<grid, kExtractThreads, 0, ctx->compute_stream>>>(jt.d_slots, reinterpret_cast<const uint8_t*>(jt.d_slots) + uint64_t(jt.capacity) * jt.slot_u4 * sizeof(uint4), jt.capacity, jt.slot_u4, static_cast<uint4*>(dev_rows_out),
                                                           capacity_rows, d_cnt);
  CU(ctx, cudaGetLastError());
  CU(ctx, cudaMemcpyAsync(h_cnt, d_cnt, 8, cudaMemcpyDeviceToHost, ctx->compute_stream));
  CU(ctx, cudaStreamSynchronize(ctx->compute_stream));
  if (*h_cnt > capacity_rows)
    return ctx->fail(PGF_ERR_INVALID_ARGUMENT, "join table holds %llu rows, export buffer has room for %llu", *h_cnt, (unsigned long long)capacity_rows);
  *rows_out = *h_cnt;
  return PGF_OK;
}

pgf_status join_from_fragments(pgf_ctx* ctx, const JoinTable& like, const void* dev_rows, uint64_t stride_bytes,
                               const uint64_t* counts, uint32_t nfragments, uint64_t* table_out) {
  std::lock_guard<std::mutex> g(ctx->mu);
  CU(ctx, cudaSetDevice(ctx->device));
  uint64_t total = 0;
  const uint64_t row_bytes = uint64_t(like.slot_u4) * sizeof(uint4);
  for (uint32_t f = 0; f < nfragments; ++f) {
    if (counts[f] * row_bytes > stride_bytes) return ctx->fail(PGF_ERR_INVALID_ARGUMENT, "fragment %u overruns its stride", f);
    total += counts[f];
  }
  uint64_t cap = 1024;
  while (cap < total * 2) cap <<= 1;
  if (cap > (1ull << 31)) return ctx->fail(PGF_ERR_NOT_ELIGIBLE, "join build side too large for one table");
  JoinTable jt = like;
  jt.capacity = uint32_t(cap);
  jt.rows = total;
  const uint64_t slot_bytes = cap * row_bytes, tag_bytes = cap + 16;
  jt.d_slots = static_cast<uint4*>(ctx->join_alloc(slot_bytes + tag_bytes, &jt.alloc_bytes));
  if (!jt.d_slots) return ctx->fail(PGF_ERR_OUT_OF_MEMORY, "cannot allocate a join table of %llu slots", (unsigned long long)cap);
  cudaError_t e = cudaMemsetAsync(reinterpret_cast<uint8_t*>(jt.d_slots) + slot_bytes, 0, tag_bytes, ctx->compute_stream);
  uint8_t* tags = reinterpret_cast<uint8_t*>(jt.d_slots) + slot_bytes;
  for (uint32_t f = 0; f < nfragments && e == cudaSuccess; ++f) {
    if (!counts[f]) continue;
    const uint4* rows = reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(dev_rows) + f * stride_bytes);
    const uint32_t grid = uint32_t(std::min<uint64_t>((counts[f] + 255) / 256, uint64_t(ctx->sm_count) * 8));
    uint32_t shift = 64;
    for (uint64_t c = cap; c > 1; c >>= 1) shift--;
    join_build_kernel<<<grid, 256, 0, ctx->compute_stream>>>(jt.d_slots, tags, jt.capacity - 1, shift, jt.slot_u4, rows, counts[f]);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->compute_stream);
  if (e != cudaSuccess) {
    cudaFree(jt.d_slots);
    return ctx->cuda_fail(e, "join_from_fragments", __FILE__, __LINE__);
  }
  const uint64_t id = ctx->next_handle++;
  ctx->joins[id] = jt;
  *table_out = id;
  return PGF_OK;
}

namespace {
pgf_status xchg_reserve(pgf_ctx* ctx, size_t bytes) {
  if (ctx->d_xchg_cap >= bytes) return PGF_OK;
  CU(ctx, cudaStreamSynchronize(ctx->compute_stream));
  if (ctx->d_xchg) cudaFree(ctx->d_xchg);
  ctx->d_xchg = nullptr;
  ctx->d_xchg_cap = 0;
  const size_t want = align_up(bytes + bytes / 4, 1 << 16);
  void* p = nullptr;
  if (cudaMalloc(&p, want) != cudaSuccess) {
    cudaGetLastError();
    return ctx->fail(PGF_ERR_OUT_OF_MEMORY, "cannot allocate %zu bytes of exchange scratch", want);
  }
  ctx->d_xchg = static_cast<uint8_t*>(p);
  ctx->d_xchg_cap = want;
  return PGF_OK;
}

// hash table over dense rows (caller holds ctx->mu); `rows` may be freed once the stream has passed the build
pgf_status table_from_rows(pgf_ctx* ctx, const JoinTable& like, const uint4* rows, uint64_t nrows, JoinTable* out) {
  uint64_t cap = 1024;
  while (cap < nrows * 2) cap <<= 1;
  if (cap > (1ull << 31)) return ctx->fail(PGF_ERR_NOT_ELIGIBLE, "join build side too large for one table");
  JoinTable jt = like;
  jt.d_rows = nullptr;
  jt.rows_alloc_bytes = 0;
  jt.capacity = uint32_t(cap);
  jt.rows = nrows;
  const uint64_t slot_bytes = cap * jt.slot_u4 * sizeof(uint4), tag_bytes = cap + 16;
  jt.d_slots = static_cast<uint4*>(ctx->join_alloc(slot_bytes + tag_bytes, &jt.alloc_bytes));
  if (!jt.d_slots) return ctx->fail(PGF_ERR_OUT_OF_MEMORY, "cannot allocate a join table of %llu slots", (unsigned long long)cap);
  cudaError_t e = cudaMemsetAsync(reinterpret_cast<uint8_t*>(jt.d_slots) + slot_bytes, 0, tag_bytes, ctx->compute_stream);
  if (e == cudaSuccess && nrows) {
    uint32_t shift = 64;
    for (uint64_t c = cap; c > 1; c >>= 1) shift--;
    const uint32_t grid = uint32_t(std::min<uint64_t>((nrows + 255) / 256, uint64_t(ctx->sm_count) * 8));
    join_build_kernel<<<grid, 256, 0, ctx->compute_stream>>>(jt.d_slots, reinterpret_cast<uint8_t*>(jt.d_slots) + slot_bytes, jt.capacity - 1, shift,
                                                              jt.slot_u4, rows, nrows);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) {
    ctx->join_free(jt.d_slots, jt.alloc_bytes);
    return ctx->cuda_fail(e, "table_from_rows", __FILE__, __LINE__);
  }
  *out = jt;
  return PGF_OK;
}
}  // namespace

// AggregateExec Partial -> all-gather -> Final in one call (the collectives stay behind the C ABI).
pgf_status pipeline_run_sharded(pgf_ctx* ctx, const pgf_pipeline* plan, uint64_t max_groups, pgf_result** out) {
  NvtxRange nvtx_("pgf:pipeline_run_sharded");
  if (ctx->comm_world == 1) return pipeline_run(ctx, plan, false, nullptr, 0, nullptr, false, out);
  if (plan->sink != PGF_SINK_AGGREGATE) return ctx->fail(PGF_ERR_INVALID_ARGUMENT, "pgf_pipeline_run_sharded runs aggregate sinks");
  uint64_t state_bytes = 0;
  PGF_TRY(pgf_partial_state_bytes(plan, max_groups ? max_groups : 1, &state_bytes));
  state_bytes = align_up(state_bytes, 16);
  {
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    PGF_TRY(xchg_reserve(ctx, state_bytes * (size_t(ctx->comm_world) + 1)));
  }
  uint8_t* mine = ctx->d_xchg;
  uint8_t* all = ctx->d_xchg + state_bytes;
  PGF_TRY(pipeline_run_partial_async(ctx, plan, mine, state_bytes));
  {
    std::lock_guard<std::mutex> g(ctx->mu);
    PGF_TRY(comm_all_gather(ctx, mine, all, state_bytes));
  }
  return pipeline_merge(ctx, plan, all, state_bytes, uint32_t(ctx->comm_world), true, out);
}

// Join exchange over NCCL: broadcast (all-gather of every rank's rows) or hash partition (all-to-all).
pgf_status join_exchange(pgf_ctx* ctx, uint64_t handle, uint32_t mode, uint64_t* out_handle, uint64_t* nvlink_bytes) {
  NvtxRange nvtx_("pgf:join_exchange");
  std::lock_guard<std::mutex> g(ctx->mu);
  CU(ctx, cudaSetDevice(ctx->device));
  auto it = ctx->joins.find(handle);
  if (it == ctx->joins.end()) return ctx->fail(PGF_ERR_UNKNOWN_HANDLE, "unknown join table / row set %llu", (unsigned long long)handle);
  const JoinTable src = it->second;
  const uint32_t world = uint32_t(ctx->comm_world), rank = uint32_t(ctx->comm_rank);
  if (world > kMaxRanks) return ctx->fail(PGF_ERR_INVALID_ARGUMENT, "at most %u ranks", kMaxRanks);
  const bool partition = (mode & 3u) == PGF_XCHG_PARTITION, rows_only = (mode & PGF_XCHG_ROWS_ONLY) != 0;
  const uint64_t row_bytes = uint64_t(src.slot_u4) * sizeof(uint4);
  cudaStream_t st = ctx->compute_stream;
  PhaseTrace trace(st);
  struct Buf {
    pgf_ctx* ctx;
    void* p = nullptr;
    size_t bytes = 0;
    ~Buf() { if (p) ctx->join_free(p, bytes); }
    bool alloc(size_t n) { p = ctx->join_alloc(n ? n : 16, &bytes); return p != nullptr; }
    void* release() { void* q = p; p = nullptr; return q; }
  };
  // 1. this rank's rows, dense
  Buf exported{ctx};
  const uint4* rows = src.d_rows;
  uint64_t nrows = src.rows;
  unsigned long long* d_cnt = reinterpret_cast<unsigned long long*>(ctx->d_flags + 4);   // [kMaxRanks] counters / cursors
  unsigned long long* h_cnt = reinterpret_cast<unsigned long long*>(ctx->h_flags + 4);
  if (!rows) {
    if (!exported.alloc(nrows * row_bytes)) return ctx->fail(PGF_ERR_OUT_OF_MEMORY, "exchange: row buffer");
    CU(ctx, cudaMemsetAsync(d_cnt, 0, 8, st));
    const uint32_t grid = uint32_t(std::min<uint64_t>((uint64_t(src.capacity) + kExtractThreads - 1) / kExtractThreads, uint64_t(ctx->sm_count) * 8));
    join_export_kernel<<<grid, kExtractThreads, 0, st>>>(src.d_slots, reinterpret_cast<const uint8_t*>(src.d_slots) + uint64_t(src.capacity) * src.slot_u4 * sizeof(uint4), src.capacity,
                                                         src.slot_u4, static_cast<uint4*>(exported.p), nrows, d_cnt);
    CU(ctx, cudaGetLastError());
    rows = static_cast<const uint4*>(exported.p);
  }
  uint64_t sent = 0;
  Buf recv{ctx};
  uint64_t total = 0;
  const uint32_t emu_world = (mode >> 8) & 0xFFu, emu_rank = (mode >> 16) & 0xFFu;
  if (world == 1 && partition && emu_world > 1) {
    // single-process emulation of one rank of an emu_world-way partition (tests on one GPU): the same count /
    // scatter kernels, then this rank keeps the segment it would have been sent by the only contributor
    if (emu_world > kMaxRanks || emu_rank >= emu_world) return ctx->fail(PGF_ERR_INVALID_ARGUMENT, "bad emulated rank / world");
    CU(ctx, cudaMemsetAsync(d_cnt, 0, 8 * kMaxRanks, st));
    const uint32_t grid = uint32_t(std::min<uint64_t>(std::max<uint64_t>((nrows + 255) / 256, 1), uint64_t(ctx->sm_count) * 4));
    if (nrows) partition_count_kernel<<<grid, 256, 0, st>>>(rows, nrows, src.slot_u4, emu_world, d_cnt);
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaMemcpyAsync(h_cnt, d_cnt, 8 * kMaxRanks, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaStreamSynchronize(st));
    std::vector<uint64_t> cnt(emu_world), off(emu_world);
    uint64_t o = 0;
    for (uint32_t p = 0; p < emu_world; ++p) { cnt[p] = h_cnt[p]; off[p] = o; o += cnt[p]; }
    if (o != nrows) return ctx->fail(PGF_ERR_STATE, "partition count lost rows");
    Buf sendbuf{ctx};
    if (!sendbuf.alloc(nrows * row_bytes) || !recv.alloc(cnt[emu_rank] * row_bytes)) return ctx->fail(PGF_ERR_OUT_OF_MEMORY, "exchange: partition buffers");
    for (uint32_t p = 0; p < emu_world; ++p) h_cnt[p] = off[p];
    CU(ctx, cudaMemcpyAsync(d_cnt, h_cnt, 8 * emu_world, cudaMemcpyHostToDevice, st));
    if (nrows) partition_scatter_kernel<<<grid, 256, 0, st>>>(rows, nrows, src.slot_u4, emu_world, d_cnt, static_cast<uint4*>(sendbuf.p));
    CU(ctx, cudaGetLastError());
    total = cnt[emu_rank];
    CU(ctx, cudaMemcpyAsync(recv.p, static_cast<uint8_t*>(sendbuf.p) + off[emu_rank] * row_bytes, total * row_bytes, cudaMemcpyDeviceToDevice, st));
    CU(ctx, cudaStreamSynchronize(st));
  } else if (world == 1) {
    if (!recv.alloc(nrows * row_bytes)) return ctx->fail(PGF_ERR_OUT_OF_MEMORY, "exchange: row buffer");
    CU(ctx, cudaMemcpyAsync(recv.p, rows, nrows * row_bytes, cudaMemcpyDeviceToDevice, st));
    total = nrows;
  } else if (!partition) {
    // ---- broadcast: every rank's row count, then an all-gather of fragments padded to the largest
    PGF_TRY(xchg_reserve(ctx, 8 * (size_t(world) + 1)));
    CU(ctx, cudaMemcpyAsync(ctx->d_xchg, &nrows, 8, cudaMemcpyHostToDevice, st));
    PGF_TRY(comm_all_gather(ctx, ctx->d_xchg, ctx->d_xchg + 8, 8));
    std::vector<uint64_t> counts(world);
    CU(ctx, cudaMemcpyAsync(counts.data(), ctx->d_xchg + 8, 8 * world, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaStreamSynchronize(st));
    uint64_t mx = 0;
    for (uint64_t c : counts) { mx = std::max(mx, c); total += c; }
    Buf padded{ctx}, gathered{ctx};
    if (!padded.alloc(mx * row_bytes) || !gathered.alloc(mx * row_bytes * world) || !recv.alloc(total * row_bytes))
      return ctx->fail(PGF_ERR_OUT_OF_MEMORY, "exchange: broadcast buffers");
    CU(ctx, cudaMemcpyAsync(padded.p, rows, nrows * row_bytes, cudaMemcpyDeviceToDevice, st));
    PGF_TRY(comm_all_gather(ctx, padded.p, gathered.p, mx * row_bytes));
    uint64_t off = 0;
    for (uint32_t r = 0; r < world; ++r) {   // compact the fragments (rank order)
      if (counts[r]) CU(ctx, cudaMemcpyAsync(static_cast<uint8_t*>(recv.p) + off, static_cast<uint8_t*>(gathered.p) + uint64_t(r) * mx * row_bytes,
                                             counts[r] * row_bytes, cudaMemcpyDeviceToDevice, st));
      off += counts[r] * row_bytes;
    }
    sent = mx * row_bytes * (world - 1);
    CU(ctx, cudaStreamSynchronize(st));   // padded / gathered go back to the cache
  } else {
    // ---- hash partition: count per owner, scatter into owner order, all-to-all
    CU(ctx, cudaMemsetAsync(d_cnt, 0, 8 * kMaxRanks, st));
    const uint32_t grid = uint32_t(std::min<uint64_t>(std::max<uint64_t>((nrows + 255) / 256, 1), uint64_t(ctx->sm_count) * 4));
    if (nrows) partition_count_kernel<<<grid, 256, 0, st>>>(rows, nrows, src.slot_u4, world, d_cnt);
    CU(ctx, cudaGetLastError());
    PGF_TRY(xchg_reserve(ctx, 8 * kMaxRanks * (size_t(world) + 1)));
    CU(ctx, cudaMemcpyAsync(ctx->d_xchg, d_cnt, 8 * kMaxRanks, cudaMemcpyDeviceToDevice, st));
    PGF_TRY(comm_all_gather(ctx, ctx->d_xchg, ctx->d_xchg + 8 * kMaxRanks, 8 * kMaxRanks));
    std::vector<uint64_t> matrix(size_t(world) * kMaxRanks);   // matrix[src][dst]
    CU(ctx, cudaMemcpyAsync(matrix.data(), ctx->d_xchg + 8 * kMaxRanks, 8 * kMaxRanks * world, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaStreamSynchronize(st));
    std::vector<uint64_t> soff(world), sbytes(world), roff(world), rbytes(world);
    uint64_t so = 0, ro = 0;
    for (uint32_t p = 0; p < world; ++p) {
      soff[p] = so; sbytes[p] = matrix[size_t(rank) * kMaxRanks + p] * row_bytes; so += sbytes[p];
      roff[p] = ro; rbytes[p] = matrix[size_t(p) * kMaxRanks + rank] * row_bytes; ro += rbytes[p];
      if (p != rank) sent += sbytes[p];
    }
    total = ro / row_bytes;
    Buf sendbuf{ctx};
    if (!sendbuf.alloc(so) || !recv.alloc(ro)) return ctx->fail(PGF_ERR_OUT_OF_MEMORY, "exchange: partition buffers");
    for (uint32_t p = 0; p < world; ++p) h_cnt[p] = soff[p] / row_bytes;   // cursors start at the owners' offsets
    CU(ctx, cudaMemcpyAsync(d_cnt, h_cnt, 8 * world, cudaMemcpyHostToDevice, st));
    if (nrows) partition_scatter_kernel<<<grid, 256, 0, st>>>(rows, nrows, src.slot_u4, world, d_cnt, static_cast<uint4*>(sendbuf.p));
    CU(ctx, cudaGetLastError());
    PGF_TRY(comm_all_to_all_v(ctx, sendbuf.p, soff.data(), sbytes.data(), recv.p, roff.data(), rbytes.data()));
    CU(ctx, cudaStreamSynchronize(st));   // the send buffer goes back to the cache
  }
  trace.mark(partition ? "exchange: partition + all-to-all" : "exchange: broadcast");
  // 2. the output: a row set, or a hash table over the received rows
  JoinTable outjt = src;
  outjt.d_slots = nullptr;
  outjt.alloc_bytes = 0;
  outjt.capacity = 0;
  outjt.d_rows = nullptr;
  outjt.rows_alloc_bytes = 0;
  outjt.rows = total;
  if (rows_only) {
    outjt.rows_alloc_bytes = recv.bytes;
    outjt.d_rows = static_cast<uint4*>(recv.release());
  } else {
    PGF_TRY(table_from_rows(ctx, src, static_cast<const uint4*>(recv.p), total, &outjt));
    CU(ctx, cudaStreamSynchronize(st));   // recv goes back to the cache
  }
  CU(ctx, cudaStreamSynchronize(st));   // every temporary of the exchange is idle before it returns to the cache
  trace.mark("exchange: table from rows");
  const uint64_t id = ctx->next_handle++;
  ctx->joins[id] = outjt;
  *out_handle = id;
  if (nvlink_bytes) *nvlink_bytes = sent;
  return PGF_OK;
}

}  // namespace pgf
