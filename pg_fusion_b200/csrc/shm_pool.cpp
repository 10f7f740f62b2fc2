// Shared-memory runtime-filter pool interop.
//
// pg_fusion keeps its runtime Bloom filters in a fixed-slot pool inside PostgreSQL shared
// memory (runtime_filter/src/pool.rs): the worker allocates a slot for a (session, scan) target,
// builds the filter, publishes it Ready, and every backend scanning that target probes the
// bits before it encodes a tuple.  This file speaks that binary protocol from the GPU library:
// the filter is built in HBM by the fused pipeline, its words are copied into the slot of the
// pool and the lifecycle word is flipped to Ready with the same atomic transitions, so the
// backends (unchanged reference code) keep rejecting rows before they are ever encoded.
//
// Layout (all #[repr(C)], native endian; pool.rs:152-217):
//   PoolHeader  56 B : magic u64 | version u32 | slot_count u32 | bit_count u64 | hash_count u32 |
//                      reserved u32 | seed u64 | word_count u64 | region_size u64
//   PoolSlot    48 B : state u32 | refs u32 | generation u64 | session_epoch u64 | scan_id u64 |
//                      output_column u32 | key_type u32 | lifecycle u64
//   bits             : slot_count x word_count u64, slot i at bits_base + i * word_count
// Lifecycle word = (generation << 2) | state  (shared.rs:7-9,400-416).
#include <atomic>
#include <cstring>

#include "context.hpp"   // cuda_runtime.h first: bloom_device.cuh uses its __host__ __device__ markers
#include "bloom_device.cuh"

namespace {

constexpr uint64_t kPoolMagic = 0x5047465552465031ull;  // pool.rs:12
constexpr uint32_t kPoolVersion = 1;                    // pool.rs:14
constexpr uint32_t kSlotFree = 0, kSlotAllocated = 1, kSlotRetiring = 2;  // pool.rs:16-18
constexpr uint64_t kMaxGeneration = ~0ull >> 2;         // shared.rs:9

struct PoolHeader {
  uint64_t magic;
  uint32_t version, slot_count;
  uint64_t bit_count;
  uint32_t hash_count, reserved0;
  uint64_t seed, word_count, region_size;
};
struct PoolSlot {
  std::atomic<uint32_t> state, refs;
  std::atomic<uint64_t> generation, session_epoch, scan_id;
  std::atomic<uint32_t> output_column, key_type;
  std::atomic<uint64_t> lifecycle;
};
static_assert(sizeof(PoolHeader) == 56 && sizeof(PoolSlot) == 48, "pool structs are part of the shared-memory format");
static_assert(std::atomic<uint64_t>::is_always_lock_free && std::atomic<uint32_t>::is_always_lock_free, "shm atomics must be lock free");

struct Pool {
  PoolHeader* header;
  PoolSlot* slots;
  std::atomic<uint64_t>* bits;
  uint32_t slot_count;
  uint64_t word_count;
};

bool layout_of(uint32_t slot_count, const pgf_bloom_params& p, uint64_t* size, uint64_t* bits_off) {
  // ComputedLayout::new (pool.rs:191-208): header, slots, bits; every part is 8-byte aligned
  const unsigned __int128 words = (unsigned __int128)slot_count * p.word_count;
  const unsigned __int128 total = (unsigned __int128)sizeof(PoolHeader) + (unsigned __int128)slot_count * sizeof(PoolSlot) + words * 8;
  if (total > (unsigned __int128)1 << 62) return false;
  *bits_off = sizeof(PoolHeader) + uint64_t(slot_count) * sizeof(PoolSlot);
  *size = uint64_t(total);
  return true;
}

pgf_status open_pool(void* base, uint64_t len, uint32_t slot_count, const pgf_bloom_params* p, bool check_header, Pool* out) {
  if (!base || !p) return PGF_ERR_INVALID_ARGUMENT;             // NullBase
  if (reinterpret_cast<uintptr_t>(base) & 7) return PGF_ERR_INVALID_ARGUMENT;  // Misaligned
  uint64_t size, bits_off;
  if (!layout_of(slot_count, *p, &size, &bits_off)) return PGF_ERR_BLOOM_TOO_MANY_BITS;  // LayoutOverflow
  if (len < size) return PGF_ERR_BLOOM_INSUFFICIENT_WORDS;      // TooSmall
  uint8_t* b = static_cast<uint8_t*>(base);
  out->header = reinterpret_cast<PoolHeader*>(b);
  out->slots = reinterpret_cast<PoolSlot*>(b + sizeof(PoolHeader));
  out->bits = reinterpret_cast<std::atomic<uint64_t>*>(b + bits_off);
  out->slot_count = slot_count;
  out->word_count = p->word_count;
  if (check_header) {  // RuntimeFilterPool::attach (pool.rs:318-349)
    const PoolHeader& h = *out->header;
    if (h.magic != kPoolMagic || h.version != kPoolVersion) return PGF_ERR_STATE;
    if (h.slot_count != slot_count || h.bit_count != p->bit_count || h.hash_count != uint32_t(p->hash_count) ||
        h.seed != p->seed || h.word_count != p->word_count || h.region_size != size)
      return PGF_ERR_STATE;  // ConfigMismatch
  }
  return PGF_OK;
}

pgf_status transition_build(PoolSlot& s, uint64_t generation, uint64_t next_state) {  // shared.rs:376-397
  uint64_t expected = (generation << 2) | PGF_RF_BUILDING;
  return s.lifecycle.compare_exchange_strong(expected, (generation << 2) | next_state, std::memory_order_acq_rel, std::memory_order_acquire)
             ? PGF_OK : PGF_ERR_LIFECYCLE_INVALID_TRANSITION;
}

void release_ref(Pool& pool, uint32_t slot_index) {  // pool.rs:527-556
  PoolSlot& s = pool.slots[slot_index];
  const uint32_t old = s.refs.fetch_sub(1, std::memory_order_acq_rel);
  if (old == 1 && s.state.load(std::memory_order_acquire) == kSlotRetiring) {
    const uint64_t generation = s.generation.load(std::memory_order_acquire);
    const uint64_t word = s.lifecycle.load(std::memory_order_acquire);
    if ((word & 3) == PGF_RF_READY) {  // retire_ready_after_quiescence: last reference, no probe inside a bit read
      uint64_t expected = (generation << 2) | PGF_RF_READY;
      s.lifecycle.compare_exchange_strong(expected, (generation << 2) | PGF_RF_DISABLED, std::memory_order_acq_rel, std::memory_order_acquire);
    } else if ((word & 3) == PGF_RF_BUILDING) {
      transition_build(s, generation, PGF_RF_DISABLED);
    }
    s.session_epoch.store(0, std::memory_order_release);
    s.scan_id.store(0, std::memory_order_release);
    s.output_column.store(0, std::memory_order_release);
    s.key_type.store(0, std::memory_order_release);
    s.generation.store(0, std::memory_order_release);
    s.state.store(kSlotFree, std::memory_order_release);
  }
}

}  // namespace

extern "C" {

pgf_status pgf_shm_pool_layout(uint32_t slot_count, const pgf_bloom_params* params, uint64_t* size_out, uint64_t* align_out) {
  if (!params || !size_out) return PGF_ERR_INVALID_ARGUMENT;
  uint64_t size, bits_off;
  if (!layout_of(slot_count, *params, &size, &bits_off)) return PGF_ERR_BLOOM_TOO_MANY_BITS;
  *size_out = size;
  if (align_out) *align_out = 8;
  return PGF_OK;
}

pgf_status pgf_shm_pool_init(void* base, uint64_t len, uint32_t slot_count, const pgf_bloom_params* params) {
  Pool pool;
  PGF_TRY(open_pool(base, len, slot_count, params, false, &pool));
  uint64_t size = 0, bits_off = 0;
  if (!layout_of(slot_count, *params, &size, &bits_off)) return PGF_ERR_BLOOM_TOO_MANY_BITS;
  std::memset(base, 0, size);  // slots Free, refs 0, lifecycle Free generation 0, bits clear
  *pool.header = PoolHeader{kPoolMagic, kPoolVersion, slot_count, params->bit_count, uint32_t(params->hash_count), 0,
                            params->seed, params->word_count, size};
  return PGF_OK;
}

pgf_status pgf_shm_pool_attach_check(void* base, uint64_t len, uint32_t slot_count, const pgf_bloom_params* params) {
  Pool pool;
  return open_pool(base, len, slot_count, params, true, &pool);
}

pgf_status pgf_shm_pool_allocate_build(void* base, uint64_t len, uint32_t slot_count, const pgf_bloom_params* params,
                                       const pgf_rf_target* target, int32_t* slot_index_out, uint64_t* generation_out) {
  if (!target || !slot_index_out || !generation_out) return PGF_ERR_INVALID_ARGUMENT;
  if (target->key_type < 1 || target->key_type > 3) return PGF_ERR_INVALID_ARGUMENT;
  Pool pool;
  PGF_TRY(open_pool(base, len, slot_count, params, true, &pool));
  *slot_index_out = -1;  // pool exhausted: a soft miss, the query runs without a runtime filter
  for (uint32_t i = 0; i < slot_count; ++i) {
    PoolSlot& s = pool.slots[i];
    uint32_t expect = kSlotFree;
    if (!s.state.compare_exchange_strong(expect, kSlotAllocated, std::memory_order_acq_rel, std::memory_order_acquire)) continue;
    // The owner's reference.  The reference stores 1 here (pool.rs:401); a backend whose lookup_probes pinned
    // the slot between the CAS above and this line would lose its pin to that store and later drive the
    // count below the owner's.  An increment keeps such a transient pin (a Free slot has refs == 0, so the
    // result is the same 1 when nobody interferes) and is indistinguishable on the wire.
    s.refs.fetch_add(1, std::memory_order_acq_rel);
    s.session_epoch.store(target->session_epoch, std::memory_order_release);
    s.scan_id.store(target->scan_id, std::memory_order_release);
    s.output_column.store(target->output_column, std::memory_order_release);
    s.key_type.store(target->key_type, std::memory_order_release);
    // try_acquire_builder (shared.rs:159-198): Free | Disabled -> Building(generation + 1), bits cleared
    for (;;) {
      uint64_t cur = s.lifecycle.load(std::memory_order_acquire);
      const uint64_t state = cur & 3, generation = cur >> 2;
      pgf_status err = PGF_OK;
      if (state == PGF_RF_BUILDING || state == PGF_RF_READY) err = PGF_ERR_LIFECYCLE_BUSY;
      else if (generation + 1 > kMaxGeneration) err = PGF_ERR_LIFECYCLE_GENERATION_EXHAUSTED;
      if (err) {
        s.refs.fetch_sub(1, std::memory_order_acq_rel);  // (the reference stores 0, pool.rs:423; see above)
        s.state.store(kSlotFree, std::memory_order_release);
        return err;
      }
      if (s.lifecycle.compare_exchange_strong(cur, ((generation + 1) << 2) | PGF_RF_BUILDING, std::memory_order_acq_rel,
                                              std::memory_order_acquire)) {
        std::atomic<uint64_t>* bits = pool.bits + uint64_t(i) * pool.word_count;
        for (uint64_t w = 0; w < pool.word_count; ++w) bits[w].store(0, std::memory_order_relaxed);
        s.generation.store(generation + 1, std::memory_order_release);
        *slot_index_out = int32_t(i);
        *generation_out = generation + 1;
        return PGF_OK;
      }
    }
  }
  return PGF_OK;
}

pgf_status pgf_shm_pool_publish_words(void* base, uint64_t len, uint32_t slot_count, const pgf_bloom_params* params,
                                      int32_t slot_index, uint64_t generation, const uint64_t* words, uint64_t nwords) {
  Pool pool;
  PGF_TRY(open_pool(base, len, slot_count, params, true, &pool));
  if (slot_index < 0 || uint32_t(slot_index) >= slot_count || !words || nwords != pool.word_count) return PGF_ERR_INVALID_ARGUMENT;
  PoolSlot& s = pool.slots[slot_index];
  // insert_hash only touches the bits of a Building slot of this generation (pool.rs:478-490)
  if (s.lifecycle.load(std::memory_order_acquire) != ((generation << 2) | PGF_RF_BUILDING)) return PGF_ERR_LIFECYCLE_INVALID_TRANSITION;
  std::atomic<uint64_t>* bits = pool.bits + uint64_t(slot_index) * pool.word_count;
  for (uint64_t w = 0; w < nwords; ++w)
    if (words[w]) bits[w].fetch_or(words[w], std::memory_order_relaxed);
  return transition_build(s, generation, PGF_RF_READY);  // publish_build: release-publishes the bits
}

pgf_status pgf_shm_pool_disable_build(void* base, uint64_t len, uint32_t slot_count, const pgf_bloom_params* params,
                                      int32_t slot_index, uint64_t generation) {
  Pool pool;
  PGF_TRY(open_pool(base, len, slot_count, params, true, &pool));
  if (slot_index < 0 || uint32_t(slot_index) >= slot_count) return PGF_ERR_INVALID_ARGUMENT;
  return transition_build(pool.slots[slot_index], generation, PGF_RF_DISABLED);
}

pgf_status pgf_shm_pool_release_owner(void* base, uint64_t len, uint32_t slot_count, const pgf_bloom_params* params,
                                      int32_t slot_index) {
  Pool pool;
  PGF_TRY(open_pool(base, len, slot_count, params, true, &pool));
  if (slot_index < 0 || uint32_t(slot_index) >= slot_count) return PGF_ERR_INVALID_ARGUMENT;
  uint32_t expect = kSlotAllocated;  // release_owner (pool.rs:516-525)
  pool.slots[slot_index].state.compare_exchange_strong(expect, kSlotRetiring, std::memory_order_acq_rel, std::memory_order_acquire);
  release_ref(pool, uint32_t(slot_index));
  return PGF_OK;
}

/* ---- probe side (what a backend does; pool.rs:432-476, shared.rs:350-374).  The worker never needs it in
 * production -- the backends run the reference's own code -- but it completes the protocol so the
 * reference's pool tests can be replayed against this implementation. */
pgf_status pgf_shm_pool_lookup_probes(void* base, uint64_t len, uint32_t slot_count, const pgf_bloom_params* params,
                                      uint64_t session_epoch, uint64_t scan_id, pgf_pool_probe* out, uint32_t max_probes,
                                      uint32_t* nprobes_out) {
  if (!nprobes_out || (!out && max_probes)) return PGF_ERR_INVALID_ARGUMENT;
  Pool pool;
  PGF_TRY(open_pool(base, len, slot_count, params, true, &pool));
  uint32_t n = 0;
  for (uint32_t i = 0; i < slot_count && n < max_probes; ++i) {
    PoolSlot& s = pool.slots[i];
    if (s.state.load(std::memory_order_acquire) != kSlotAllocated) continue;
    s.refs.fetch_add(1, std::memory_order_acq_rel);   // pin first, then re-check (the owner may be retiring)
    const bool matches = s.state.load(std::memory_order_acquire) == kSlotAllocated &&
                         s.session_epoch.load(std::memory_order_acquire) == session_epoch &&
                         s.scan_id.load(std::memory_order_acquire) == scan_id;
    const uint32_t kt = s.key_type.load(std::memory_order_acquire);
    if (!matches || kt < 1 || kt > 3) {
      release_ref(pool, i);
      continue;
    }
    out[n++] = pgf_pool_probe{int32_t(i), kt, s.generation.load(std::memory_order_acquire),
                              s.output_column.load(std::memory_order_acquire), 0};
  }
  *nprobes_out = n;
  return PGF_OK;
}

pgf_status pgf_shm_pool_probe_decide(void* base, uint64_t len, uint32_t slot_count, const pgf_bloom_params* params,
                                     int32_t slot_index, uint64_t generation, int32_t key_is_null, int64_t key,
                                     int32_t* decision_out) {
  if (!decision_out) return PGF_ERR_INVALID_ARGUMENT;
  Pool pool;
  PGF_TRY(open_pool(base, len, slot_count, params, true, &pool));
  if (slot_index < 0 || uint32_t(slot_index) >= slot_count) return PGF_ERR_INVALID_ARGUMENT;
  const uint64_t word = pool.slots[slot_index].lifecycle.load(std::memory_order_acquire);
  if (word != ((generation << 2) | PGF_RF_READY)) {   // Free / Building / Disabled / another generation never reject
    *decision_out = PGF_PASS_UNFILTERED;
    return PGF_OK;
  }
  if (key_is_null) {                                   // decision_for_null: a NULL key cannot match an inner join
    *decision_out = PGF_DEFINITELY_ABSENT;
    return PGF_OK;
  }
  const std::atomic<uint64_t>* bits = pool.bits + uint64_t(slot_index) * pool.word_count;
  const uint64_t h1 = pgf::splitmix64(uint64_t(key) ^ params->seed);           // bloom.rs:250-255
  const uint64_t h2 = pgf::splitmix64(h1 ^ pgf::kBloomSalt) | 1ull;
  uint64_t v = h1;
  *decision_out = PGF_MAYBE_PRESENT;
  for (uint64_t i = 0; i < params->hash_count; ++i, v += h2) {
    const uint64_t bit = v % params->bit_count;
    if (((bits[bit >> 6].load(std::memory_order_relaxed) >> (bit & 63)) & 1ull) == 0) {
      *decision_out = PGF_DEFINITELY_ABSENT;
      break;
    }
  }
  return PGF_OK;
}

pgf_status pgf_shm_pool_release_probe(void* base, uint64_t len, uint32_t slot_count, const pgf_bloom_params* params,
                                      int32_t slot_index) {
  Pool pool;
  PGF_TRY(open_pool(base, len, slot_count, params, true, &pool));
  if (slot_index < 0 || uint32_t(slot_index) >= slot_count) return PGF_ERR_INVALID_ARGUMENT;
  release_ref(pool, uint32_t(slot_index));
  return PGF_OK;
}

pgf_status pgf_bloom_publish_to_pool(pgf_ctx* ctx, uint64_t bloom, void* base, uint64_t len, uint32_t slot_count,
                                     int32_t slot_index, uint64_t generation) {
  if (!ctx) return PGF_ERR_INVALID_ARGUMENT;
  if (ctx->sticky) return ctx->sticky;
  auto it = ctx->blooms.find(bloom);
  if (it == ctx->blooms.end()) return ctx->fail(PGF_ERR_UNKNOWN_HANDLE, "unknown bloom filter %llu", (unsigned long long)bloom);
  const pgf::BloomSlot& b = it->second;
  std::vector<uint64_t> words(b.params.word_count);
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaStreamSynchronize(ctx->compute_stream));
  CU(ctx, cudaMemcpy(words.data(), b.d_words, words.size() * 8, cudaMemcpyDeviceToHost));
  pgf_status st = pgf_shm_pool_publish_words(base, len, slot_count, &b.params, slot_index, generation, words.data(), words.size());
  if (st) return ctx->fail(st, "cannot publish the filter into pool slot %d (generation %llu)", slot_index, (unsigned long long)generation);
  return PGF_OK;
}

}  // extern "C"
