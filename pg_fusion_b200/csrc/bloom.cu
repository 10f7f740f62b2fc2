// Runtime Bloom filter kernels (build K1 / probe K2 of SURVEY.md section 2.2).
//   build : RuntimeFilterBuildStream::insert_batch -> insert_ints -> insert_hash
//           (worker_runtime/src/runtime_filter_plan.rs:227-274,345-363;
//            runtime_filter/src/bloom.rs:222-227)
//   probe : runtime_filter_rejects_slot -> decision_for_hash / decision_for_null
//           (pg/backend_service/src/source.rs:496-532; runtime_filter/src/shared.rs:350-374)
// Keys stream with 128-bit coalesced loads (HBM bound: 8 B/key in, 1 B/key out); the bit
// array (128 KiB at the GUC defaults) is staged in shared memory when it fits, otherwise it
// stays L2 resident.
#include "bloom_device.cuh"
#include "context.hpp"
#include "layout.hpp"

namespace pgf {

namespace {

constexpr int kThreads = 512;
constexpr uint32_t kSmemWordsMax = 24 * 1024;  // 192 KiB of filter words per CTA

__device__ __forceinline__ int64_t load_key(const void* keys, int width, uint64_t i) {
  // sign-extend Int16/Int32 to i64 (runtime_filter_plan.rs:244,256,268)
  if (width == 8) return reinterpret_cast<const int64_t*>(keys)[i];
  if (width == 4) return int64_t(reinterpret_cast<const int32_t*>(keys)[i]);
  return int64_t(reinterpret_cast<const int16_t*>(keys)[i]);
}

__device__ __forceinline__ bool valid_bit(const uint8_t* validity, uint64_t i) {
  return validity == nullptr || ((validity[i >> 3] >> (i & 7)) & 1);
}

struct KeySpan {  // a contiguous run of keys: a host array copy, or one column of one page
  const void* keys;
  const uint8_t* validity;
  uint64_t n;
  uint64_t out_base;
};

struct ScanKeys {
  const uint8_t* pages;
  const PageDesc* descs;
  const LayoutClass* classes;
  uint64_t page_stride;
  uint32_t npages;
  uint32_t col;
  int32_t width;
  int32_t nullable;
};

__device__ __forceinline__ KeySpan page_span(const ScanKeys& s, uint32_t page) {
  const PageDesc d = s.descs[page];
  const LayoutClass& lc = s.classes[d.layout_class];
  const uint8_t* base = s.pages + page * s.page_stride;
  KeySpan sp;
  sp.keys = base + lc.values_off[s.col];
  sp.validity = (s.nullable && ((d.null_mask >> s.col) & 1)) ? base + lc.validity_off[s.col] : nullptr;
  sp.n = d.row_count;
  sp.out_base = d.row_base;
  return sp;
}

// ---- build -----------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) bloom_insert_array_kernel(DevBloom b, KeySpan sp, int width,
                                                                     unsigned long long* inserted) {
  unsigned long long mine = 0;
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < sp.n; i += uint64_t(gridDim.x) * blockDim.x) {
    if (!valid_bit(sp.validity, i)) continue;  // NULL keys are never inserted
    bloom_insert(b, uint64_t(load_key(sp.keys, width, i)));
    ++mine;
  }
  for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(inserted, mine);
}

__global__ void __launch_bounds__(kThreads) bloom_insert_scan_kernel(DevBloom b, ScanKeys s,
                                                                    unsigned long long* inserted) {
  unsigned long long mine = 0;
  for (uint32_t page = blockIdx.x; page < s.npages; page += gridDim.x) {
    const KeySpan sp = page_span(s, page);
    for (uint32_t i = threadIdx.x; i < sp.n; i += blockDim.x) {
      if (!valid_bit(sp.validity, i)) continue;
      bloom_insert(b, uint64_t(load_key(sp.keys, s.width, i)));
      ++mine;
    }
  }
  for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(inserted, mine);
}

// ---- probe -----------------------------------------------------------------------
// Decision per key; `ready` is the lifecycle check hoisted out of the per-key path
// (the reference re-loads the lifecycle word per key, shared.rs:351-354).
//
// The probe is bound by integer issue, not by HBM: two splitmix64 rounds are ~50 SASS
// instructions per key against 9 bytes of traffic.  Two things keep the rest cheap:
//   * K > 0 (hash_count == K and bit_count a power of two <= 2^32): bit_i only depends on
//     the low 32 bits of h1 + i*h2, the filter is read as 32-bit words (little endian:
//     u64 word b/64, bit b%64 == u32 word b/32, bit b%32) and all K words are fetched
//     before any is tested, so the K loads overlap instead of forming a dependent chain;
//   * 4 keys per thread per iteration are loaded one iteration ahead (32 KiB in flight per SM).
// K == 0 is the general form (any bit_count via the exact fastmod, early exit per bit).
template <bool SMEM, int K>
__device__ __forceinline__ bool probe_words(const DevBloom& b, const uint64_t* smem_words, uint64_t key) {
  const uint64_t h1 = splitmix64(key ^ b.seed);
  const uint64_t h2 = splitmix64(h1 ^ kBloomSalt) | 1ull;
  if constexpr (K > 0) {
    const uint32_t* w32 = reinterpret_cast<const uint32_t*>(SMEM ? smem_words : b.words);
    const uint32_t mask = uint32_t(b.bit_count - 1);
    uint32_t v = uint32_t(h1), acc = 1u;
    const uint32_t step = uint32_t(h2);
#pragma unroll
    for (int i = 0; i < K; ++i, v += step) {
      const uint32_t bit = v & mask;
      const uint32_t w = SMEM ? w32[bit >> 5] : __ldg(w32 + (bit >> 5));
      acc &= w >> (bit & 31);
    }
    return acc & 1u;
  } else {
    const uint64_t* words = SMEM ? smem_words : b.words;
    uint64_t v = h1;
    for (uint32_t i = 0; i < b.hash_count; ++i, v += h2) {
      const uint64_t bit = bloom_reduce(b, v);
      const uint64_t w = SMEM ? words[bit >> 6] : __ldg(words + (bit >> 6));
      if (((w >> (bit & 63)) & 1ull) == 0) return false;
    }
    return true;
  }
}

template <bool SMEM, int K>
__device__ __forceinline__ uint8_t decide(const DevBloom& b, const uint64_t* smem_words, bool ready,
                                          bool valid, int64_t key) {
  if (!ready) return PGF_PASS_UNFILTERED;
  if (!valid) return PGF_DEFINITELY_ABSENT;  // decision_for_null, shared.rs:367-374
  return probe_words<SMEM, K>(b, smem_words, uint64_t(key)) ? PGF_MAYBE_PRESENT : PGF_DEFINITELY_ABSENT;
}

template <bool SMEM>
__device__ __forceinline__ void stage_words(const DevBloom& b, uint64_t* smem_words, uint32_t nwords) {
  if (!SMEM) return;
  // 128-bit coalesced copy of the bit array into shared memory
  const uint4* src = reinterpret_cast<const uint4*>(b.words);
  uint4* dst = reinterpret_cast<uint4*>(smem_words);
  for (uint32_t i = threadIdx.x; i < nwords / 2; i += blockDim.x) dst[i] = __ldg(src + i);
  if ((nwords & 1) && threadIdx.x == 0) smem_words[nwords - 1] = b.words[nwords - 1];
  __syncthreads();
}

struct ProbeOut {
  uint8_t* decisions;
  unsigned long long* rejected;
  unsigned long long* unfiltered;
};

__device__ __forceinline__ longlong2 ldg_stream_i64x2(const longlong2* p) {
  longlong2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.s64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
  return v;
}

__device__ __forceinline__ void finish_counts(unsigned long long rej, unsigned long long unf, const ProbeOut& out) {
  for (int o = 16; o; o >>= 1) {
    rej += __shfl_xor_sync(0xffffffffu, rej, o);
    unf += __shfl_xor_sync(0xffffffffu, unf, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (rej) atomicAdd(out.rejected, rej);
    if (unf) atomicAdd(out.unfiltered, unf);
  }
}

// A block of 4 * kProbeThreads consecutive Int64 keys without NULLs: thread t owns the key
// pairs (2t, 2t+1) of the two halves of the block.
constexpr int kProbeThreads = 1024;
constexpr uint32_t kBlockKeys = 4u * kProbeThreads;
struct KeyQuad {
  longlong2 a, b;
};
__device__ __forceinline__ KeyQuad load_quad(const void* keys, uint32_t n, uint32_t base) {
  // rows past n are not loaded; pairs start at even rows so a pair is in range iff its first row is
  KeyQuad q{{0, 0}, {0, 0}};
  const longlong2* k2 = reinterpret_cast<const longlong2*>(keys);
  const uint32_t r0 = base + 2u * threadIdx.x, r1 = r0 + 2u * kProbeThreads;
  if (r0 + 1 < n) q.a = ldg_stream_i64x2(k2 + (r0 >> 1));
  else if (r0 < n) q.a.x = reinterpret_cast<const int64_t*>(keys)[r0];
  if (r1 + 1 < n) q.b = ldg_stream_i64x2(k2 + (r1 >> 1));
  else if (r1 < n) q.b.x = reinterpret_cast<const int64_t*>(keys)[r1];
  return q;
}
template <bool SMEM, int K>
__device__ __forceinline__ void probe_quad(const DevBloom& b, const uint64_t* smem_words, bool ready, const KeyQuad& q,
                                           uint32_t n, uint32_t base, uint8_t* out, unsigned long long& rej,
                                           unsigned long long& unf) {
  const uint8_t d0 = decide<SMEM, K>(b, smem_words, ready, true, q.a.x), d1 = decide<SMEM, K>(b, smem_words, ready, true, q.a.y);
  const uint8_t d2 = decide<SMEM, K>(b, smem_words, ready, true, q.b.x), d3 = decide<SMEM, K>(b, smem_words, ready, true, q.b.y);
  const uint32_t r0 = base + 2u * threadIdx.x, r1 = r0 + 2u * kProbeThreads;
  const bool aligned = (reinterpret_cast<uintptr_t>(out) & 1) == 0;
  auto put = [&](uint32_t r, uint8_t x, uint8_t y) {
    if (r + 1 < n) {
      if (aligned) *reinterpret_cast<uchar2*>(out + r) = make_uchar2(x, y);
      else { out[r] = x; out[r + 1] = y; }
      rej += (x == PGF_DEFINITELY_ABSENT) + (y == PGF_DEFINITELY_ABSENT);
      unf += (x == PGF_PASS_UNFILTERED) + (y == PGF_PASS_UNFILTERED);
    } else if (r < n) {
      out[r] = x;
      rej += x == PGF_DEFINITELY_ABSENT;
      unf += x == PGF_PASS_UNFILTERED;
    }
  };
  put(r0, d0, d1);
  put(r1, d2, d3);
}

// Host key arrays (copied to a device buffer).
template <bool SMEM, int K>
__global__ void __launch_bounds__(kProbeThreads, 1) bloom_probe_array_kernel(DevBloom b, KeySpan sp, int width, bool ready,
                                                                            uint32_t nwords, ProbeOut out) {
  extern __shared__ __align__(16) uint64_t smem_words[];
  stage_words<SMEM>(b, smem_words, nwords);
  unsigned long long rej = 0, unf = 0;
  if (width == 8 && sp.validity == nullptr && (reinterpret_cast<uintptr_t>(sp.keys) & 15) == 0 && sp.n < 0xFFFF0000ull) {
    const uint32_t n = uint32_t(sp.n), nblocks = (n + kBlockKeys - 1) / kBlockKeys;
    uint32_t blk = blockIdx.x;
    KeyQuad q = blk < nblocks ? load_quad(sp.keys, n, blk * kBlockKeys) : KeyQuad{};
    for (; blk < nblocks; blk += gridDim.x) {
      const uint32_t nxt = blk + gridDim.x;
      const KeyQuad qn = nxt < nblocks ? load_quad(sp.keys, n, nxt * kBlockKeys) : KeyQuad{};
      probe_quad<SMEM, K>(b, smem_words, ready, q, n, blk * kBlockKeys, out.decisions + sp.out_base, rej, unf);
      q = qn;
    }
  } else {
    const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < sp.n; i += stride) {
      const uint8_t d = decide<SMEM, K>(b, smem_words, ready, valid_bit(sp.validity, i), load_key(sp.keys, width, i));
      out.decisions[sp.out_base + i] = d;
      rej += d == PGF_DEFINITELY_ABSENT;
      unf += d == PGF_PASS_UNFILTERED;
    }
  }
  finish_counts(rej, unf, out);
}

// One key column of a scan's pages.  Work item = one block of kBlockKeys rows of one page.
template <bool SMEM, int K>
__global__ void __launch_bounds__(kProbeThreads, 1) bloom_probe_scan_kernel(DevBloom b, ScanKeys s, bool ready,
                                                                           uint32_t nwords, uint32_t blocks_per_page,
                                                                           ProbeOut out) {
  extern __shared__ __align__(16) uint64_t smem_words[];
  stage_words<SMEM>(b, smem_words, nwords);
  unsigned long long rej = 0, unf = 0;
  const uint32_t nitems = s.npages * blocks_per_page;
  if (s.width == 8) {
    auto fetch = [&](uint32_t item, KeySpan& sp, uint32_t& base) -> KeyQuad {
      sp = page_span(s, item / blocks_per_page);
      base = (item % blocks_per_page) * kBlockKeys;
      // pages with NULLs take the row-at-a-time path below
      return sp.validity == nullptr ? load_quad(sp.keys, uint32_t(sp.n), base) : KeyQuad{};
    };
    uint32_t item = blockIdx.x;
    KeySpan sp{}, spn{};
    uint32_t base = 0, basen = 0;
    KeyQuad q{};
    if (item < nitems) q = fetch(item, sp, base);
    for (; item < nitems; item += gridDim.x) {
      const uint32_t nxt = item + gridDim.x;
      KeyQuad qn{};
      if (nxt < nitems) qn = fetch(nxt, spn, basen);
      if (sp.validity == nullptr) {
        probe_quad<SMEM, K>(b, smem_words, ready, q, uint32_t(sp.n), base, out.decisions + sp.out_base, rej, unf);
      } else {
        const uint32_t end = uint32_t(sp.n) < base + kBlockKeys ? uint32_t(sp.n) : base + kBlockKeys;
        for (uint32_t i = base + threadIdx.x; i < end; i += blockDim.x) {
          const uint8_t d = decide<SMEM, K>(b, smem_words, ready, valid_bit(sp.validity, i), load_key(sp.keys, 8, i));
          out.decisions[sp.out_base + i] = d;
          rej += d == PGF_DEFINITELY_ABSENT;
          unf += d == PGF_PASS_UNFILTERED;
        }
      }
      q = qn; sp = spn; base = basen;
    }
  } else {
    for (uint32_t item = blockIdx.x; item < nitems; item += gridDim.x) {
      const KeySpan sp = page_span(s, item / blocks_per_page);
      const uint32_t base = (item % blocks_per_page) * kBlockKeys;
      const uint32_t end = uint32_t(sp.n) < base + kBlockKeys ? uint32_t(sp.n) : base + kBlockKeys;
#pragma unroll 4
      for (uint32_t i = base + threadIdx.x; i < end; i += blockDim.x) {
        const uint8_t d = decide<SMEM, K>(b, smem_words, ready, valid_bit(sp.validity, i), load_key(sp.keys, s.width, i));
        out.decisions[sp.out_base + i] = d;
        rej += d == PGF_DEFINITELY_ABSENT;
        unf += d == PGF_PASS_UNFILTERED;
      }
    }
  }
  finish_counts(rej, unf, out);
}

__global__ void bloom_or_kernel(uint64_t* dst, const uint64_t* src, uint64_t nwords, uint32_t narrays) {
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < nwords; i += uint64_t(gridDim.x) * blockDim.x) {
    uint64_t v = dst[i];
    for (uint32_t a = 0; a < narrays; ++a) v |= src[a * nwords + i];
    dst[i] = v;
  }
}

__global__ void bloom_popcount_kernel(const uint64_t* words, uint64_t nwords, unsigned long long* out) {
  unsigned long long mine = 0;
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < nwords; i += uint64_t(gridDim.x) * blockDim.x) mine += __popcll(words[i]);
  for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(out, mine);
}

int key_width_of(int type) {
  switch (type) {
    case PGF_T_INT16: return 2;
    case PGF_T_INT32: return 4;
    case PGF_T_INT64: return 8;
    default: return 0;  // key_type_for, runtime_filter_plan.rs:113-120
  }
}

uint32_t grid_for(pgf_ctx* ctx, uint64_t work_items, uint32_t per_sm) {
  const uint64_t full = uint64_t(ctx->sm_count) * per_sm;
  const uint64_t need = (work_items + kThreads - 1) / kThreads;
  return uint32_t(need < full ? (need ? need : 1) : full);
}

pgf_status make_scan_keys(pgf_ctx* ctx, Scan& s, uint32_t col, ScanKeys* out) {
  if (col >= s.schema.size()) return ctx->fail(PGF_ERR_INVALID_ARGUMENT, "column %u out of range", col);
  const int w = key_width_of(s.schema[col].type_tag);
  if (!w) return ctx->fail(PGF_ERR_NOT_ELIGIBLE, "runtime filter keys must be Int16/Int32/Int64");
  PGF_TRY(scan_sync_descs(ctx, s));
  out->pages = s.d_pages;
  out->descs = s.d_descs;
  out->classes = s.d_classes;
  out->page_stride = ctx->page_size;
  out->npages = uint32_t(s.npages);
  out->col = col;
  out->width = w;
  out->nullable = s.schema[col].nullable;
  return PGF_OK;
}

// copies a host array to a temporary device buffer on the compute stream
struct DevTemp {
  void* p = nullptr;
  ~DevTemp() { if (p) cudaFree(p); }
};

}  // namespace

pgf_status bloom_make_dev(const pgf_bloom_params& p, uint64_t* d_words, DevBloom* out) {
  out->words = d_words;
  out->bit_count = p.bit_count;
  out->seed = p.seed;
  out->hash_count = uint32_t(p.hash_count > 0xFFFFFFFFull ? 0xFFFFFFFFull : p.hash_count);
  out->pow2 = (p.bit_count & (p.bit_count - 1)) == 0;
  if (out->pow2 && p.bit_count >= 32 && p.bit_count <= (1ull << 32)) out->pow2 = 2;  // 32-bit word addressing
  // M = floor((2^128 - 1) / d) + 1
  const unsigned __int128 all = ~(unsigned __int128)0;
  const unsigned __int128 m = all / p.bit_count + 1;
  out->m_lo = uint64_t(m);
  out->m_hi = uint64_t(m >> 64);
  return PGF_OK;
}

pgf_status bloom_insert_host_keys(pgf_ctx* ctx, BloomSlot& b, const void* keys, int32_t key_width,
                                  const uint8_t* validity, uint64_t n, uint64_t* inserted) {
  CU(ctx, cudaSetDevice(ctx->device));
  if (n == 0) {
    if (inserted) *inserted = 0;
    return PGF_OK;
  }
  DevTemp dk, dv;
  CU(ctx, cudaMalloc(&dk.p, n * uint64_t(key_width)));
  CU(ctx, cudaMemcpyAsync(dk.p, keys, n * uint64_t(key_width), cudaMemcpyHostToDevice, ctx->compute_stream));
  if (validity) {
    CU(ctx, cudaMalloc(&dv.p, (n + 7) / 8));
    CU(ctx, cudaMemcpyAsync(dv.p, validity, (n + 7) / 8, cudaMemcpyHostToDevice, ctx->compute_stream));
  }
  unsigned long long* d_cnt = reinterpret_cast<unsigned long long*>(ctx->d_flags + 8);
  unsigned long long* h_cnt = reinterpret_cast<unsigned long long*>(ctx->h_flags + 8);
  CU(ctx, cudaMemsetAsync(d_cnt, 0, 8, ctx->compute_stream));
  KeySpan sp{dk.p, static_cast<const uint8_t*>(dv.p), n, 0};
  CU(ctx, cudaEventRecord(ctx->ev_a, ctx->compute_stream));
  bloom_insert_array_kernel<<<grid_for(ctx, n, 4), kThreads, 0, ctx->compute_stream>>>(b.dev, sp, key_width, d_cnt);
  CU(ctx, cudaGetLastError());
  CU(ctx, cudaEventRecord(ctx->ev_b, ctx->compute_stream));
  CU(ctx, cudaMemcpyAsync(h_cnt, d_cnt, 8, cudaMemcpyDeviceToHost, ctx->compute_stream));
  CU(ctx, cudaStreamSynchronize(ctx->compute_stream));
  if (inserted) *inserted = *h_cnt;
  CU(ctx, cudaEventElapsedTime(&ctx->last_kernel_ms, ctx->ev_a, ctx->ev_b));
  return PGF_OK;
}

pgf_status bloom_insert_scan(pgf_ctx* ctx, BloomSlot& b, Scan& s, uint32_t col, uint64_t* inserted) {
  CU(ctx, cudaSetDevice(ctx->device));
  ScanKeys sk;
  PGF_TRY(make_scan_keys(ctx, s, col, &sk));
  unsigned long long* d_cnt = reinterpret_cast<unsigned long long*>(ctx->d_flags + 8);
  unsigned long long* h_cnt = reinterpret_cast<unsigned long long*>(ctx->h_flags + 8);
  CU(ctx, cudaMemsetAsync(d_cnt, 0, 8, ctx->compute_stream));
  CU(ctx, cudaEventRecord(ctx->ev_a, ctx->compute_stream));
  if (s.npages) {
    const uint32_t grid = uint32_t(s.npages < uint64_t(ctx->sm_count) * 4 ? s.npages : uint64_t(ctx->sm_count) * 4);
    bloom_insert_scan_kernel<<<grid, kThreads, 0, ctx->compute_stream>>>(b.dev, sk, d_cnt);
    CU(ctx, cudaGetLastError());
  }
  CU(ctx, cudaEventRecord(ctx->ev_b, ctx->compute_stream));
  CU(ctx, cudaMemcpyAsync(h_cnt, d_cnt, 8, cudaMemcpyDeviceToHost, ctx->compute_stream));
  CU(ctx, cudaStreamSynchronize(ctx->compute_stream));
  if (inserted) *inserted = *h_cnt;
  CU(ctx, cudaEventElapsedTime(&ctx->last_kernel_ms, ctx->ev_a, ctx->ev_b));
  return PGF_OK;
}

namespace {
template <class Launch>
pgf_status run_probe(pgf_ctx* ctx, BloomSlot& b, uint64_t n, uint8_t* decisions, pgf_probe_stats* stats,
                     Launch launch) {
  DevTemp dd;
  CU(ctx, cudaMalloc(&dd.p, n ? n : 1));
  unsigned long long* d_cnt = reinterpret_cast<unsigned long long*>(ctx->d_flags + 8);
  unsigned long long* h_cnt = reinterpret_cast<unsigned long long*>(ctx->h_flags + 8);
  CU(ctx, cudaMemsetAsync(d_cnt, 0, 16, ctx->compute_stream));
  ProbeOut out{static_cast<uint8_t*>(dd.p), d_cnt, d_cnt + 1};
  const uint32_t nwords = uint32_t(b.params.word_count);
  const bool smem = b.params.word_count <= kSmemWordsMax;
  CU(ctx, cudaEventRecord(ctx->ev_a, ctx->compute_stream));
  if (n) PGF_TRY(launch(out, smem, nwords));
  CU(ctx, cudaEventRecord(ctx->ev_b, ctx->compute_stream));
  CU(ctx, cudaMemcpyAsync(h_cnt, d_cnt, 16, cudaMemcpyDeviceToHost, ctx->compute_stream));
  if (n) CU(ctx, cudaMemcpyAsync(decisions, dd.p, n, cudaMemcpyDeviceToHost, ctx->compute_stream));
  CU(ctx, cudaStreamSynchronize(ctx->compute_stream));
  CU(ctx, cudaEventElapsedTime(&ctx->last_kernel_ms, ctx->ev_a, ctx->ev_b));
  if (stats) {
    stats->probe_rows = n;
    stats->rejected_rows = h_cnt[0];
    stats->pass_unfiltered = h_cnt[1];
  }
  return PGF_OK;
}
}  // namespace

namespace {
// hash_count with an unrolled instantiation (pow2 bit_count <= 2^32), else 0 = general form
int probe_k(const BloomSlot& b) {
  const uint64_t bits = b.params.bit_count;
  const bool pow2 = (bits & (bits - 1)) == 0 && bits <= (1ull << 32) && bits >= 32;
  return (pow2 && b.params.hash_count >= 1 && b.params.hash_count <= 8) ? int(b.params.hash_count) : 0;
}

template <bool SMEM, int K>
cudaError_t launch_probe_array(uint32_t grid, size_t bytes, cudaStream_t st, const DevBloom& b, const KeySpan& sp, int width,
                               bool ready, uint32_t nwords, const ProbeOut& out) {
  auto kernel = bloom_probe_array_kernel<SMEM, K>;
  if (SMEM) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
    if (e != cudaSuccess) return e;
  }
  kernel<<<grid, kProbeThreads, bytes, st>>>(b, sp, width, ready, nwords, out);
  return cudaGetLastError();
}
template <bool SMEM, int K>
cudaError_t launch_probe_scan(uint32_t grid, size_t bytes, cudaStream_t st, const DevBloom& b, const ScanKeys& sk, bool ready,
                              uint32_t nwords, uint32_t blocks_per_page, const ProbeOut& out) {
  auto kernel = bloom_probe_scan_kernel<SMEM, K>;
  if (SMEM) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
    if (e != cudaSuccess) return e;
  }
  kernel<<<grid, kProbeThreads, bytes, st>>>(b, sk, ready, nwords, blocks_per_page, out);
  return cudaGetLastError();
}

#define PGF_PROBE_DISPATCH(FN, SMEM_FLAG, K, ...)                                            \
  [&]() -> cudaError_t {                                                                     \
    switch (K) {                                                                             \
      case 1: return SMEM_FLAG ? FN<true, 1>(__VA_ARGS__) : FN<false, 1>(__VA_ARGS__);       \
      case 2: return SMEM_FLAG ? FN<true, 2>(__VA_ARGS__) : FN<false, 2>(__VA_ARGS__);       \
      case 3: return SMEM_FLAG ? FN<true, 3>(__VA_ARGS__) : FN<false, 3>(__VA_ARGS__);       \
      case 4: return SMEM_FLAG ? FN<true, 4>(__VA_ARGS__) : FN<false, 4>(__VA_ARGS__);       \
      case 5: return SMEM_FLAG ? FN<true, 5>(__VA_ARGS__) : FN<false, 5>(__VA_ARGS__);       \
      case 6: return SMEM_FLAG ? FN<true, 6>(__VA_ARGS__) : FN<false, 6>(__VA_ARGS__);       \
      case 7: return SMEM_FLAG ? FN<true, 7>(__VA_ARGS__) : FN<false, 7>(__VA_ARGS__);       \
      case 8: return SMEM_FLAG ? FN<true, 8>(__VA_ARGS__) : FN<false, 8>(__VA_ARGS__);       \
      default: return SMEM_FLAG ? FN<true, 0>(__VA_ARGS__) : FN<false, 0>(__VA_ARGS__);      \
    }                                                                                        \
  }()
}  // namespace

pgf_status bloom_probe_host_keys(pgf_ctx* ctx, BloomSlot& b, bool ready, const void* keys, int32_t key_width,
                                 const uint8_t* validity, uint64_t n, uint8_t* decisions, pgf_probe_stats* stats) {
  CU(ctx, cudaSetDevice(ctx->device));
  DevTemp dk, dv;
  if (n) {
    CU(ctx, cudaMalloc(&dk.p, n * uint64_t(key_width)));
    CU(ctx, cudaMemcpyAsync(dk.p, keys, n * uint64_t(key_width), cudaMemcpyHostToDevice, ctx->compute_stream));
    if (validity) {
      CU(ctx, cudaMalloc(&dv.p, (n + 7) / 8));
      CU(ctx, cudaMemcpyAsync(dv.p, validity, (n + 7) / 8, cudaMemcpyHostToDevice, ctx->compute_stream));
    }
  }
  KeySpan sp{dk.p, static_cast<const uint8_t*>(dv.p), n, 0};
  return run_probe(ctx, b, n, decisions, stats, [&](ProbeOut out, bool smem, uint32_t nwords) -> pgf_status {
    const uint64_t blocks = (n + kBlockKeys - 1) / kBlockKeys;
    const uint32_t grid = uint32_t(blocks < uint64_t(ctx->sm_count) ? (blocks ? blocks : 1) : uint64_t(ctx->sm_count));
    const size_t bytes = smem ? size_t(nwords) * 8 : 0;
    const int k = probe_k(b);
    CU(ctx, PGF_PROBE_DISPATCH(launch_probe_array, smem, k, grid, bytes, ctx->compute_stream, b.dev, sp, key_width, ready, nwords, out));
    return PGF_OK;
  });
}

pgf_status bloom_probe_scan(pgf_ctx* ctx, BloomSlot& b, bool ready, Scan& s, uint32_t col, uint8_t* decisions,
                            pgf_probe_stats* stats) {
  CU(ctx, cudaSetDevice(ctx->device));
  ScanKeys sk;
  PGF_TRY(make_scan_keys(ctx, s, col, &sk));
  return run_probe(ctx, b, s.rows, decisions, stats, [&](ProbeOut out, bool smem, uint32_t nwords) -> pgf_status {
    const uint32_t max_rows = s.max_page_rows ? s.max_page_rows : 1;
    const uint32_t blocks_per_page = (max_rows + kBlockKeys - 1) / kBlockKeys;
    const uint64_t items = s.npages * uint64_t(blocks_per_page);
    if (items > 0xFFFFFFF0ull) return ctx->fail(PGF_ERR_NOT_ELIGIBLE, "scan too large for 32-bit probe items");
    const uint32_t grid = uint32_t(items < uint64_t(ctx->sm_count) ? (items ? items : 1) : uint64_t(ctx->sm_count));
    const size_t bytes = smem ? size_t(nwords) * 8 : 0;
    const int k = probe_k(b);
    CU(ctx, PGF_PROBE_DISPATCH(launch_probe_scan, smem, k, grid, bytes, ctx->compute_stream, b.dev, sk, ready, nwords, blocks_per_page, out));
    return PGF_OK;
  });
}

pgf_status bloom_count_bits(pgf_ctx* ctx, BloomSlot& b) {
  CU(ctx, cudaSetDevice(ctx->device));
  unsigned long long* d_cnt = reinterpret_cast<unsigned long long*>(ctx->d_flags + 8);
  unsigned long long* h_cnt = reinterpret_cast<unsigned long long*>(ctx->h_flags + 8);
  CU(ctx, cudaMemsetAsync(d_cnt, 0, 8, ctx->compute_stream));
  const uint64_t nwords = b.params.word_count;
  const uint32_t grid = uint32_t((nwords + 255) / 256 < 1184 ? (nwords + 255) / 256 : 1184);
  bloom_popcount_kernel<<<grid ? grid : 1, 256, 0, ctx->compute_stream>>>(b.d_words, nwords, d_cnt);
  CU(ctx, cudaGetLastError());
  CU(ctx, cudaMemcpyAsync(h_cnt, d_cnt, 8, cudaMemcpyDeviceToHost, ctx->compute_stream));
  CU(ctx, cudaStreamSynchronize(ctx->compute_stream));
  b.set_bits = *h_cnt;
  return PGF_OK;
}

// dst[i] |= src[a * nwords + i] for a < narrays (the caller holds the device)
pgf_status bloom_or_strided(pgf_ctx* ctx, uint64_t* dst, const uint64_t* src, uint64_t nwords, uint32_t narrays, uint32_t grid) {
  bloom_or_kernel<<<grid ? grid : 1, 256, 0, ctx->compute_stream>>>(dst, src, nwords, narrays);
  CU(ctx, cudaGetLastError());
  return PGF_OK;
}

pgf_status bloom_or_device(pgf_ctx* ctx, BloomSlot& b, const void* dev_words, uint64_t nwords, uint32_t narrays) {
  CU(ctx, cudaSetDevice(ctx->device));
  const uint32_t grid = uint32_t((nwords + 255) / 256 < 1184 ? (nwords + 255) / 256 : 1184);
  bloom_or_kernel<<<grid ? grid : 1, 256, 0, ctx->compute_stream>>>(b.d_words, static_cast<const uint64_t*>(dev_words), nwords, narrays);
  CU(ctx, cudaGetLastError());
  return PGF_OK;
}

}  // namespace pgf
