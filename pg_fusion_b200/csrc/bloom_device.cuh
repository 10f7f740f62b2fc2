// Device-side runtime Bloom filter: same hash functions and bit layout as pg_fusion's
// runtime_filter crate, so bit arrays are interchangeable with the CPU implementation.
//   h1 = splitmix64(hash ^ seed); h2 = splitmix64(h1 ^ SALT) | 1;
//   bit_i = (h1 + i*h2 mod 2^64) % bit_count          (runtime_filter/src/bloom.rs:250-255)
//   word = bit / 64, mask = 1 << (bit % 64)            (runtime_filter/src/bloom.rs:243-247)
#pragma once
#include <cstdint>

#include "device_types.cuh"

namespace pgf {

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t v) {  // bloom.rs:293-299
  v += 0x9E3779B97F4A7C15ull;
  v = (v ^ (v >> 30)) * 0xBF58476D1CE4E5B9ull;
  v = (v ^ (v >> 27)) * 0x94D049BB133111EBull;
  return v ^ (v >> 31);
}

constexpr uint64_t kBloomSalt = 0xD1B54A32D192ED03ull;  // bloom.rs:10

#ifdef __CUDACC__
// Predicated read-only loads.  A load under `if (p)` becomes a branch, and loads behind
// different branches are issued and waited for one after the other; a predicated instruction
// keeps the code straight-line so independent loads of several rows overlap.
__device__ __forceinline__ uint32_t ldg_u32_if(const uint32_t* p, bool pred, uint32_t otherwise) {
  uint32_t v = otherwise;
  asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p ld.global.nc.u32 %0, [%1];\n\t}" : "+r"(v) : "l"(p), "r"(uint32_t(pred)));
  return v;
}
__device__ __forceinline__ uint4 ldg_u128_if(const uint4* p, bool pred) {
  uint4 v = make_uint4(0, 0, 0, 0);
  asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t@p ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];\n\t}"
      : "+r"(v.x), "+r"(v.y), "+r"(v.z), "+r"(v.w) : "l"(p), "r"(uint32_t(pred)));
  return v;
}
#endif

#ifdef __CUDACC__
// Exact a % d for 64-bit operands without a divide (Lemire, Kaser, Kurz 2019):
// M = ceil(2^128 / d) (128 bit), lowbits = M * a mod 2^128, result = (lowbits * d) >> 128.
__device__ __forceinline__ uint64_t fastmod_u64(uint64_t a, uint64_t m_lo, uint64_t m_hi, uint64_t d) {
  const uint64_t lb_lo = m_lo * a;
  const uint64_t lb_hi = __umul64hi(m_lo, a) + m_hi * a;
  const uint64_t t_hi = __umul64hi(lb_lo, d);
  const uint64_t p_lo = lb_hi * d;
  const uint64_t p_hi = __umul64hi(lb_hi, d);
  return p_hi + ((p_lo + t_hi) < p_lo ? 1ull : 0ull);
}

__device__ __forceinline__ uint64_t bloom_reduce(const DevBloom& b, uint64_t v) {
  return b.pow2 ? (v & (b.bit_count - 1)) : fastmod_u64(v, b.m_lo, b.m_hi, b.bit_count);
}

// b.pow2 == 2: bit_count is a power of two in [32, 2^32].  Then bit_i only depends on the low
// 32 bits of h1 + i*h2 and the filter can be addressed as 32-bit words (little endian: u64
// word b/64, bit b%64 is u32 word b/32, bit b%32) -- same bits, half the register traffic.

// AtomicBloomRef::insert_hash (bloom.rs:222-227).  The word is read first and the atomic
// skipped when the bit is already set: bits are only ever set during a build, so a stale
// "set" observation is always valid, and saturated filters stop generating atomics.  The
// reads of up to four bit positions are issued together (independent L2 round trips).
__device__ __forceinline__ void bloom_insert(const DevBloom& b, uint64_t hash) {
  const uint64_t h1 = splitmix64(hash ^ b.seed);
  const uint64_t h2 = splitmix64(h1 ^ kBloomSalt) | 1ull;
  if (b.pow2 == 2) {
    unsigned int* w32 = reinterpret_cast<unsigned int*>(b.words);
    const uint32_t mask = uint32_t(b.bit_count - 1), step = uint32_t(h2);
    uint32_t v = uint32_t(h1);
    for (uint32_t i = 0; i < b.hash_count; i += 4, v += 4u * step) {
      uint32_t bit[4], cur[4];
#pragma unroll
      for (uint32_t q = 0; q < 4; ++q) {
        bit[q] = (v + q * step) & mask;
        cur[q] = ~0u;
        if (i + q < b.hash_count) cur[q] = *reinterpret_cast<volatile unsigned int*>(w32 + (bit[q] >> 5));
      }
#pragma unroll
      for (uint32_t q = 0; q < 4; ++q) {
        const uint32_t m = 1u << (bit[q] & 31);
        if ((cur[q] & m) == 0) atomicOr(w32 + (bit[q] >> 5), m);
      }
    }
    return;
  }
  uint64_t v = h1;
  for (uint32_t i = 0; i < b.hash_count; ++i, v += h2) {
    const uint64_t bit = bloom_reduce(b, v);
    const uint64_t mask = 1ull << (bit & 63);
    unsigned long long* w = reinterpret_cast<unsigned long long*>(b.words + (bit >> 6));
    if ((*reinterpret_cast<volatile unsigned long long*>(w) & mask) == 0) atomicOr(w, mask);
  }
}

// AtomicBloomRef::might_contain_hash (bloom.rs:233-241).  The decision is the AND of the k
// bits whatever the order of evaluation: the first bit is tested alone (a sparse filter
// rejects most absent keys there), the remaining ones are fetched three at a time so their
// L1/L2 latencies overlap instead of forming a dependent chain.
__device__ __forceinline__ bool bloom_contains(const DevBloom& b, uint64_t hash) {
  const uint64_t h1 = splitmix64(hash ^ b.seed);
  const uint64_t h2 = splitmix64(h1 ^ kBloomSalt) | 1ull;
  if (b.pow2 == 2) {
    const uint32_t* w32 = reinterpret_cast<const uint32_t*>(b.words);
    const uint32_t mask = uint32_t(b.bit_count - 1), step = uint32_t(h2);
    uint32_t v = uint32_t(h1);
    uint32_t bit = v & mask;
    if (((__ldg(w32 + (bit >> 5)) >> (bit & 31)) & 1u) == 0) return false;
    for (uint32_t i = 1; i < b.hash_count; i += 3) {
      const uint32_t b0 = (v + step) & mask, b1 = (v + 2u * step) & mask, b2 = (v + 3u * step) & mask;
      uint32_t x0 = __ldg(w32 + (b0 >> 5)) >> (b0 & 31), x1 = ~0u, x2 = ~0u;
      if (i + 1 < b.hash_count) x1 = __ldg(w32 + (b1 >> 5)) >> (b1 & 31);
      if (i + 2 < b.hash_count) x2 = __ldg(w32 + (b2 >> 5)) >> (b2 & 31);
      if (((x0 & x1 & x2) & 1u) == 0) return false;
      v += 3u * step;
    }
    return true;
  }
  uint64_t v = h1;
  for (uint32_t i = 0; i < b.hash_count; ++i, v += h2) {
    const uint64_t bit = bloom_reduce(b, v);
    if (((__ldg(b.words + (bit >> 6)) >> (bit & 63)) & 1ull) == 0) return false;
  }
  return true;
}

// N keys at once (the fused pipeline handles N rows per thread): the hashes and the loads of
// all keys are interleaved so their latencies overlap.  keep[] says which keys are live on
// entry and holds the decisions on return.
template <uint32_t N>
__device__ __forceinline__ void bloom_contains_n(const DevBloom& b, const uint64_t (&key)[N], bool (&keep)[N]) {
  if (b.pow2 != 2) {
#pragma unroll
    for (uint32_t q = 0; q < N; ++q)
      if (keep[q]) keep[q] = bloom_contains(b, key[q]);
    return;
  }
  const uint32_t* w32 = reinterpret_cast<const uint32_t*>(b.words);
  const uint32_t mask = uint32_t(b.bit_count - 1);
  uint64_t h1[N];
  uint32_t v[N], x[N], step[N];
  // first bit of every key
#pragma unroll
  for (uint32_t q = 0; q < N; ++q) {
    h1[q] = splitmix64(key[q] ^ b.seed);
    v[q] = uint32_t(h1[q]);
    const uint32_t bit = v[q] & mask;
    x[q] = ldg_u32_if(w32 + (bit >> 5), keep[q], 0u) >> (bit & 31);
  }
  // second hash while the loads are in flight
#pragma unroll
  for (uint32_t q = 0; q < N; ++q) step[q] = uint32_t(splitmix64(h1[q] ^ kBloomSalt)) | 1u;
  bool any = false;
#pragma unroll
  for (uint32_t q = 0; q < N; ++q) { keep[q] = keep[q] && (x[q] & 1u); any |= keep[q]; }
  for (uint32_t i = 1; i < b.hash_count && any; i += 3) {
#pragma unroll
    for (uint32_t q = 0; q < N; ++q) x[q] = ~0u;
#pragma unroll
    for (uint32_t t = 1; t <= 3; ++t) {
      if (i + t - 1 < b.hash_count) {
#pragma unroll
        for (uint32_t q = 0; q < N; ++q) {
          const uint32_t bit = (v[q] + t * step[q]) & mask;
          x[q] &= ldg_u32_if(w32 + (bit >> 5), keep[q], ~0u) >> (bit & 31);
        }
      }
    }
    any = false;
#pragma unroll
    for (uint32_t q = 0; q < N; ++q) { keep[q] = keep[q] && (x[q] & 1u); v[q] += 3u * step[q]; any |= keep[q]; }
  }
}
#endif

}  // namespace pgf
