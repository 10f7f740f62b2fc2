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

// AtomicBloomRef::insert_hash (bloom.rs:222-227).  The word is read first and the atomic
// skipped when the bit is already set: bits are only ever set during a build, so a stale
// "set" observation is always valid, and saturated filters stop generating atomics.
__device__ __forceinline__ void bloom_insert(const DevBloom& b, uint64_t hash) {
  const uint64_t h1 = splitmix64(hash ^ b.seed);
  const uint64_t h2 = splitmix64(h1 ^ kBloomSalt) | 1ull;
  uint64_t v = h1;
  for (uint32_t i = 0; i < b.hash_count; ++i, v += h2) {
    const uint64_t bit = bloom_reduce(b, v);
    const uint64_t mask = 1ull << (bit & 63);
    unsigned long long* w = reinterpret_cast<unsigned long long*>(b.words + (bit >> 6));
    if ((*reinterpret_cast<volatile unsigned long long*>(w) & mask) == 0) atomicOr(w, mask);
  }
}

// AtomicBloomRef::might_contain_hash (bloom.rs:233-241), early-out on the first clear bit.
__device__ __forceinline__ bool bloom_contains(const DevBloom& b, uint64_t hash) {
  const uint64_t h1 = splitmix64(hash ^ b.seed);
  const uint64_t h2 = splitmix64(h1 ^ kBloomSalt) | 1ull;
  uint64_t v = h1;
  for (uint32_t i = 0; i < b.hash_count; ++i, v += h2) {
    const uint64_t bit = bloom_reduce(b, v);
    if (((__ldg(b.words + (bit >> 6)) >> (bit & 63)) & 1ull) == 0) return false;
  }
  return true;
}
#endif

}  // namespace pgf
