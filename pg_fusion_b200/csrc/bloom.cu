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
template <bool SMEM>
__device__ __forceinline__ uint8_t decide(const DevBloom& b, const uint64_t* smem_words, bool ready,
                                          bool valid, int64_t key) {
  if (!ready) return PGF_PASS_UNFILTERED;
  if (!valid) return PGF_DEFINITELY_ABSENT;  // decision_for_null, shared.rs:367-374
  if (SMEM) {
    const uint64_t h1 = splitmix64(uint64_t(key) ^ b.seed);
    const uint64_t h2 = splitmix64(h1 ^ kBloomSalt) | 1ull;
    uint64_t v = h1;
    for (uint32_t i = 0; i < b.hash_count; ++i, v += h2) {
      const uint64_t bit = bloom_reduce(b, v);
      if (((smem_words[bit >> 6] >> (bit & 63)) & 1ull) == 0) return PGF_DEFINITELY_ABSENT;
    }
    return PGF_MAYBE_PRESENT;
  }
  return bloom_contains(b, uint64_t(key)) ? PGF_MAYBE_PRESENT : PGF_DEFINITELY_ABSENT;
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

// Int64 keys, no nulls: each thread handles 2 keys per 128-bit load and writes 2 bytes.
template <bool SMEM>
__global__ void __launch_bounds__(kThreads) bloom_probe_array_kernel(DevBloom b, KeySpan sp, int width, bool ready,
                                                                    uint32_t nwords, ProbeOut out) {
  extern __shared__ __align__(16) uint64_t smem_words[];
  stage_words<SMEM>(b, smem_words, nwords);
  unsigned long long rej = 0, unf = 0;
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  const uint64_t tid = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  if (width == 8 && sp.validity == nullptr && (reinterpret_cast<uintptr_t>(sp.keys) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(out.decisions + sp.out_base) & 1) == 0) {
    const longlong2* k2 = reinterpret_cast<const longlong2*>(sp.keys);
    const uint64_t pairs = sp.n / 2;
    for (uint64_t i = tid; i < pairs; i += stride) {
      const longlong2 k = __ldg(k2 + i);
      const uint8_t d0 = decide<SMEM>(b, smem_words, ready, true, k.x);
      const uint8_t d1 = decide<SMEM>(b, smem_words, ready, true, k.y);
      *reinterpret_cast<uchar2*>(out.decisions + sp.out_base + 2 * i) = make_uchar2(d0, d1);
      rej += (d0 == PGF_DEFINITELY_ABSENT) + (d1 == PGF_DEFINITELY_ABSENT);
      unf += (d0 == PGF_PASS_UNFILTERED) + (d1 == PGF_PASS_UNFILTERED);
    }
    if ((sp.n & 1) && tid == 0) {
      const uint8_t d = decide<SMEM>(b, smem_words, ready, true, load_key(sp.keys, 8, sp.n - 1));
      out.decisions[sp.out_base + sp.n - 1] = d;
      rej += d == PGF_DEFINITELY_ABSENT;
      unf += d == PGF_PASS_UNFILTERED;
    }
  } else {
    for (uint64_t i = tid; i < sp.n; i += stride) {
      const uint8_t d = decide<SMEM>(b, smem_words, ready, valid_bit(sp.validity, i), load_key(sp.keys, width, i));
      out.decisions[sp.out_base + i] = d;
      rej += d == PGF_DEFINITELY_ABSENT;
      unf += d == PGF_PASS_UNFILTERED;
    }
  }
  for (int o = 16; o; o >>= 1) {
    rej += __shfl_xor_sync(0xffffffffu, rej, o);
    unf += __shfl_xor_sync(0xffffffffu, unf, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (rej) atomicAdd(out.rejected, rej);
    if (unf) atomicAdd(out.unfiltered, unf);
  }
}

template <bool SMEM>
__global__ void __launch_bounds__(kThreads) bloom_probe_scan_kernel(DevBloom b, ScanKeys s, bool ready,
                                                                   uint32_t nwords, ProbeOut out) {
  extern __shared__ __align__(16) uint64_t smem_words[];
  stage_words<SMEM>(b, smem_words, nwords);
  unsigned long long rej = 0, unf = 0;
  for (uint32_t page = blockIdx.x; page < s.npages; page += gridDim.x) {
    const KeySpan sp = page_span(s, page);
    for (uint32_t i = threadIdx.x; i < sp.n; i += blockDim.x) {
      const uint8_t d = decide<SMEM>(b, smem_words, ready, valid_bit(sp.validity, i), load_key(sp.keys, s.width, i));
      out.decisions[sp.out_base + i] = d;
      rej += d == PGF_DEFINITELY_ABSENT;
      unf += d == PGF_PASS_UNFILTERED;
    }
  }
  for (int o = 16; o; o >>= 1) {
    rej += __shfl_xor_sync(0xffffffffu, rej, o);
    unf += __shfl_xor_sync(0xffffffffu, unf, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (rej) atomicAdd(out.rejected, rej);
    if (unf) atomicAdd(out.unfiltered, unf);
  }
}

__global__ void bloom_or_kernel(uint64_t* dst, const uint64_t* src, uint64_t nwords, uint32_t narrays) {
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < nwords; i += uint64_t(gridDim.x) * blockDim.x) {
    uint64_t v = dst[i];
    for (uint32_t a = 0; a < narrays; ++a) v |= src[a * nwords + i];
    dst[i] = v;
  }
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
    const uint32_t grid = grid_for(ctx, n / 2 + 1, 1);
    if (smem) {
      const size_t bytes = size_t(nwords) * 8;
      CU(ctx, cudaFuncSetAttribute(bloom_probe_array_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)));
      bloom_probe_array_kernel<true><<<grid, kThreads, bytes, ctx->compute_stream>>>(b.dev, sp, key_width, ready, nwords, out);
    } else {
      bloom_probe_array_kernel<false><<<grid, kThreads, 0, ctx->compute_stream>>>(b.dev, sp, key_width, ready, nwords, out);
    }
    CU(ctx, cudaGetLastError());
    return PGF_OK;
  });
}

pgf_status bloom_probe_scan(pgf_ctx* ctx, BloomSlot& b, bool ready, Scan& s, uint32_t col, uint8_t* decisions,
                            pgf_probe_stats* stats) {
  CU(ctx, cudaSetDevice(ctx->device));
  ScanKeys sk;
  PGF_TRY(make_scan_keys(ctx, s, col, &sk));
  return run_probe(ctx, b, s.rows, decisions, stats, [&](ProbeOut out, bool smem, uint32_t nwords) -> pgf_status {
    const uint32_t grid = uint32_t(s.npages < uint64_t(ctx->sm_count) ? s.npages : uint64_t(ctx->sm_count));
    if (smem) {
      const size_t bytes = size_t(nwords) * 8;
      CU(ctx, cudaFuncSetAttribute(bloom_probe_scan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)));
      bloom_probe_scan_kernel<true><<<grid, kThreads, bytes, ctx->compute_stream>>>(b.dev, sk, ready, nwords, out);
    } else {
      bloom_probe_scan_kernel<false><<<grid, kThreads, 0, ctx->compute_stream>>>(b.dev, sk, ready, nwords, out);
    }
    CU(ctx, cudaGetLastError());
    return PGF_OK;
  });
}

pgf_status bloom_or_device(pgf_ctx* ctx, BloomSlot& b, const void* dev_words, uint64_t nwords, uint32_t narrays) {
  CU(ctx, cudaSetDevice(ctx->device));
  const uint32_t grid = uint32_t((nwords + 255) / 256 < 1184 ? (nwords + 255) / 256 : 1184);
  bloom_or_kernel<<<grid ? grid : 1, 256, 0, ctx->compute_stream>>>(b.d_words, static_cast<const uint64_t*>(dev_words), nwords, narrays);
  CU(ctx, cudaGetLastError());
  return PGF_OK;
}

}  // namespace pgf
