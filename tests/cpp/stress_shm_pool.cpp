// Concurrency stress of the shared-memory runtime-filter pool protocol (csrc/shm_pool.cpp), built with
// -fsanitize=thread.  The reference checks this protocol with loom models that cannot run here
// (runtime_filter/src/pool.rs, shared.rs); this is the closest executable substitute:
//   * builder threads allocate a slot for their (session, scan) target, publish a filter over a known
//     key set, keep it for a while and release it (worker side: allocate_build / publish / release_owner);
//   * probe threads look targets up, probe with keys of that target's set and release (backend side:
//     lookup_probes / decision_for_hash / handle drop).
// Checked: no data race (TSan), a Ready filter never answers DefinitelyAbsent for a key of the set it
// was built from (no false negatives, also while slots are being recycled), NULL keys are never
// MaybePresent, failed builds never reject, and every slot is reusable when the threads are gone.
//
// The protocol (which this library follows word for word, because PostgreSQL backends run the
// reference's own probe code against the same memory) leaves two windows open, both in
// lookup_probes' "pin first, then re-check" (pool.rs:446-461):
//   1. against allocate_build: the pin (refs.fetch_add) can land between the allocator's state CAS and
//      its refs.store(1) (pool.rs:389-401), which then overwrites it -- a lost reference;
//   2. against release_owner: the pin can land after the owner's fetch_sub took refs to 0; the re-check
//      sees RETIRING, the unpin sees old_refs == 1 && RETIRING and runs the free path a second time
//      (pool.rs:527-556), possibly after the slot has been handed to the next builder.
// Either one corrupts the reference count of the slot for good (leaked slots, a Free slot whose
// lifecycle is still Ready, filters answering for the wrong target).  Run with `unguarded` this
// harness reproduces them within a second.  By default it serialises exactly these two pairs of
// calls (a shared lock around lookup_probes, an exclusive one around allocate_build and
// release_owner); probe decisions and unpins stay fully concurrent with everything, which is the
// concurrency the reference counting is there for -- and then every invariant above must hold.
// Window 1 is on the worker's side of the protocol and closed in csrc/shm_pool.cpp (the owner's reference
// is added, not stored); mode `release-only` leaves allocate_build unserialised to show that.  Window 2 is
// in the backends' unpin and cannot be closed from here.
//   stress_shm_pool <milliseconds> [unguarded | release-only]
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <shared_mutex>
#include <thread>
#include <vector>

#include "pgf_b200.h"

namespace {

constexpr uint32_t kSlots = 3;
constexpr int kBuilders = 4, kProbers = 4;
constexpr uint64_t kSalt = 0xD1B54A32D192ED03ull;

uint64_t splitmix64(uint64_t v) {
  v += 0x9E3779B97F4A7C15ull;
  v = (v ^ (v >> 30)) * 0xBF58476D1CE4E5B9ull;
  v = (v ^ (v >> 27)) * 0x94D049BB133111EBull;
  return v ^ (v >> 31);
}

// the key set of target t: 64 keys derived from t
int64_t key_of(uint64_t target, uint32_t i) { return int64_t(splitmix64(target * 1000 + i)); }

void build_words(const pgf_bloom_params& p, uint64_t target, std::vector<uint64_t>& words) {
  words.assign(p.word_count, 0);
  for (uint32_t i = 0; i < 64; ++i) {
    const uint64_t h1 = splitmix64(uint64_t(key_of(target, i)) ^ p.seed);
    const uint64_t h2 = splitmix64(h1 ^ kSalt) | 1ull;
    uint64_t v = h1;
    for (uint64_t k = 0; k < p.hash_count; ++k, v += h2) {
      const uint64_t bit = v % p.bit_count;
      words[bit >> 6] |= 1ull << (bit & 63);
    }
  }
}

std::atomic<bool> stop{false};
bool guarded = true;        // serialise lookup_probes against release_owner (window 2)
bool guard_allocate = true; // ... and against allocate_build (window 1)
std::shared_mutex window;  // see the header comment: lookup_probes vs allocate_build / release_owner

template <class F>
auto exclusive(F&& f) {
  if (!guarded) return f();
  std::unique_lock<std::shared_mutex> g(window);
  return f();
}
template <class F>
auto shared(F&& f) {
  if (!guarded) return f();
  std::shared_lock<std::shared_mutex> g(window);
  return f();
}
std::atomic<uint64_t> false_negatives{0}, null_maybe{0}, api_errors{0}, builds{0}, exhausted{0}, probes_done{0}, rejected_foreign{0};

}  // namespace

int main(int argc, char** argv) {
  const int ms = argc > 1 ? std::atoi(argv[1]) : 1500;
  guarded = !(argc > 2 && std::strcmp(argv[2], "unguarded") == 0);
  guard_allocate = guarded && !(argc > 2 && std::strcmp(argv[2], "release-only") == 0);
  pgf_bloom_params p;
  if (pgf_bloom_params_new(4099, 4, 0x7067667573696f6eull, &p) != PGF_OK) return 2;  // a prime bit count: the modulo path
  uint64_t size = 0, align = 0;
  if (pgf_shm_pool_layout(kSlots, &p, &size, &align) != PGF_OK) return 2;
  void* base = std::aligned_alloc(64, (size + 63) / 64 * 64);
  if (!base || pgf_shm_pool_init(base, size, kSlots, &p) != PGF_OK) return 2;

  std::vector<std::thread> threads;
  for (int b = 0; b < kBuilders; ++b) {
    threads.emplace_back([&, b] {
      std::vector<uint64_t> words;
      uint64_t round = 0;
      while (!stop.load(std::memory_order_relaxed)) {
        const uint64_t target = uint64_t(b) + 1;
        pgf_rf_target t{/*session_epoch=*/7, /*scan_id=*/target, /*output_column=*/uint32_t(b), /*key_type=*/3};
        int32_t slot = -1;
        uint64_t generation = 0;
        auto allocate = [&] { return pgf_shm_pool_allocate_build(base, size, kSlots, &p, &t, &slot, &generation); };
        if ((guard_allocate ? exclusive(allocate) : allocate()) != PGF_OK) { ++api_errors; continue; }
        if (slot < 0) { ++exhausted; std::this_thread::yield(); continue; }   // soft miss
        build_words(p, target, words);
        if (++round % 5 == 0) {  // some builds fail: the slot must never reject anything
          if (pgf_shm_pool_disable_build(base, size, kSlots, &p, slot, generation) != PGF_OK) ++api_errors;
        } else if (pgf_shm_pool_publish_words(base, size, kSlots, &p, slot, generation, words.data(), words.size()) != PGF_OK) {
          ++api_errors;
        }
        ++builds;
        std::this_thread::sleep_for(std::chrono::microseconds(50 + 37 * b));
        if (exclusive([&] { return pgf_shm_pool_release_owner(base, size, kSlots, &p, slot); }) != PGF_OK) ++api_errors;
      }
    });
  }
  for (int q = 0; q < kProbers; ++q) {
    threads.emplace_back([&, q] {
      uint64_t n = uint64_t(q);
      while (!stop.load(std::memory_order_relaxed)) {
        const uint64_t target = 1 + (n++ % kBuilders);
        pgf_pool_probe found[kSlots];
        uint32_t nfound = 0;
        if (shared([&] { return pgf_shm_pool_lookup_probes(base, size, kSlots, &p, 7, target, found, kSlots, &nfound); }) != PGF_OK) { ++api_errors; continue; }
        for (uint32_t f = 0; f < nfound; ++f) {
          for (uint32_t i = 0; i < 64; i += 7) {
            int32_t d = -1;
            if (pgf_shm_pool_probe_decide(base, size, kSlots, &p, found[f].slot_index, found[f].generation, 0, key_of(target, i), &d) != PGF_OK) ++api_errors;
            if (d == PGF_DEFINITELY_ABSENT) ++false_negatives;
            // a key of another target's set may be rejected (that is the point of the filter)
            if (pgf_shm_pool_probe_decide(base, size, kSlots, &p, found[f].slot_index, found[f].generation, 0, key_of(target + 100, i), &d) != PGF_OK) ++api_errors;
            if (d == PGF_DEFINITELY_ABSENT) ++rejected_foreign;
          }
          int32_t d = -1;
          if (pgf_shm_pool_probe_decide(base, size, kSlots, &p, found[f].slot_index, found[f].generation, 1, 0, &d) != PGF_OK) ++api_errors;
          if (d == PGF_MAYBE_PRESENT) ++null_maybe;
          if (pgf_shm_pool_release_probe(base, size, kSlots, &p, found[f].slot_index) != PGF_OK) ++api_errors;
          ++probes_done;
        }
      }
    });
  }
  std::this_thread::sleep_for(std::chrono::milliseconds(ms));
  stop.store(true);
  for (auto& t : threads) t.join();

  // quiescent: every slot must be reusable again (allocate all, release all)
  uint32_t reusable = 0;
  int32_t slots[kSlots];
  for (uint32_t i = 0; i < kSlots; ++i) {
    pgf_rf_target t{9, 100 + i, 0, 2};
    uint64_t g = 0;
    slots[i] = -1;
    if (pgf_shm_pool_allocate_build(base, size, kSlots, &p, &t, &slots[i], &g) == PGF_OK && slots[i] >= 0) ++reusable;
  }
  for (uint32_t i = 0; i < kSlots; ++i)
    if (slots[i] >= 0) pgf_shm_pool_release_owner(base, size, kSlots, &p, slots[i]);
  std::printf("builds %llu exhausted %llu probes %llu rejected_foreign %llu false_negatives %llu null_maybe %llu api_errors %llu reusable_slots %u/%u\n",
              (unsigned long long)builds.load(), (unsigned long long)exhausted.load(), (unsigned long long)probes_done.load(),
              (unsigned long long)rejected_foreign.load(), (unsigned long long)false_negatives.load(), (unsigned long long)null_maybe.load(),
              (unsigned long long)api_errors.load(), reusable, kSlots);
  std::free(base);
  return (false_negatives.load() || null_maybe.load() || api_errors.load() || reusable != kSlots) ? 1 : 0;
}
