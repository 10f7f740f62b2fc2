/*
 * oracle/orc_bloom.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see orc.h) for the
 * runtime Bloom filter of the reference's `runtime_filter` crate.
 *
 * Pinned by the reference's own known-answer tests (runtime_filter/src/tests.rs),
 * re-expressed in tests/test_oracle_bloom.py.
 */
#include "orc.h"

#include <math.h>
#include <string.h>

/* runtime_filter/src/bloom.rs:9-10 */
static const uint64_t HASH_GAMMA = 0x9E3779B97F4A7C15ull;
static const uint64_t HASH_SALT = 0xD1B54A32D192ED03ull;

/* runtime_filter/src/bloom.rs:293-299 -- all arithmetic wraps mod 2^64. */
uint64_t orc_splitmix64(uint64_t v) {
  v += HASH_GAMMA;
  v = (v ^ (v >> 30)) * 0xBF58476D1CE4E5B9ull;
  v = (v ^ (v >> 27)) * 0x94D049BB133111EBull;
  return v ^ (v >> 31);
}

/* runtime_filter/src/lib.rs:31-34 -- identity "hash": i64 reinterpreted as u64. */
uint64_t orc_hash_int_key(int64_t v) { return (uint64_t)v; }

/* runtime_filter/src/bloom.rs:29-48 */
int orc_bloom_params_new(uint64_t bit_count, uint64_t hash_count, uint64_t seed,
                         orc_bloom_params *out) {
  if (bit_count == 0) return ORC_BLOOM_ZERO_BIT_COUNT;
  if (hash_count == 0) return ORC_BLOOM_ZERO_HASH_COUNT;
  if (bit_count > UINT64_MAX - 63) return ORC_BLOOM_TOO_MANY_BITS;
  out->bit_count = bit_count;
  out->word_count = (bit_count + 63) / 64;
  out->hash_count = hash_count;
  out->seed = seed;
  return ORC_OK;
}

/* runtime_filter/src/bloom.rs:52-79: m = ceil(-n ln p / ln2^2), k = max(1, round(m/n ln2)). */
int orc_bloom_params_for_expected_items(uint64_t expected_items, double fpr, uint64_t seed,
                                        orc_bloom_params *out) {
  if (expected_items == 0) return ORC_BLOOM_ZERO_EXPECTED_ITEMS;
  if (!isfinite(fpr) || fpr <= 0.0 || fpr >= 1.0) return ORC_BLOOM_INVALID_FPR;
  const double expected = (double)expected_items;
  const double ln2 = 0.693147180559945309417232121458176568; /* std::f64::consts::LN_2 */
  const double bit_count_f = ceil(-(expected * log(fpr)) / (ln2 * ln2));
  if (bit_count_f > 18446744073709551615.0) return ORC_BLOOM_TOO_MANY_BITS;
  const uint64_t bit_count = (uint64_t)bit_count_f;
  double k = round(((double)bit_count / expected) * ln2); /* f64::round: half away from zero */
  if (k < 1.0) k = 1.0;
  return orc_bloom_params_new(bit_count, (uint64_t)k, seed, out);
}

/* runtime_filter/src/bloom.rs:167-178,186-197 */
int orc_bloom_attach_check(const orc_bloom_params *p, const uint64_t *bits,
                           uint64_t words_available) {
  if (bits == NULL) return ORC_BLOOM_NULL_BITS;
  if (words_available < p->word_count) return ORC_BLOOM_INSUFFICIENT_WORDS;
  return ORC_OK;
}

/* runtime_filter/src/bloom.rs:250-255 */
uint64_t orc_bloom_bit_index(const orc_bloom_params *p, uint64_t hash, uint64_t hash_index) {
  const uint64_t h1 = orc_splitmix64(hash ^ p->seed);
  const uint64_t h2 = orc_splitmix64(h1 ^ HASH_SALT) | 1ull;
  const uint64_t value = h1 + hash_index * h2;
  return value % p->bit_count;
}

/* runtime_filter/src/bloom.rs:205-209 */
void orc_bloom_clear(const orc_bloom_params *p, uint64_t *bits) {
  memset(bits, 0, (size_t)p->word_count * sizeof(uint64_t));
}

/* runtime_filter/src/bloom.rs:222-227,243-247: word = bit/64, mask = 1 << (bit%64). */
void orc_bloom_insert_hash(const orc_bloom_params *p, uint64_t *bits, uint64_t hash) {
  for (uint64_t i = 0; i < p->hash_count; ++i) {
    const uint64_t bit = orc_bloom_bit_index(p, hash, i);
    bits[bit / 64] |= 1ull << (bit % 64);
  }
}

/* runtime_filter/src/bloom.rs:233-241 */
int orc_bloom_might_contain_hash(const orc_bloom_params *p, const uint64_t *bits,
                                 uint64_t hash) {
  for (uint64_t i = 0; i < p->hash_count; ++i) {
    const uint64_t bit = orc_bloom_bit_index(p, hash, i);
    if ((bits[bit / 64] & (1ull << (bit % 64))) == 0) return 0;
  }
  return 1;
}

static inline int64_t load_key(const void *keys, int key_width, uint64_t i) {
  /* sign-extension: runtime_filter_plan.rs:244,256,268; slot_encoder/src/encoder.rs:357-365 */
  switch (key_width) {
    case 2: return (int64_t)((const int16_t *)keys)[i];
    case 4: return (int64_t)((const int32_t *)keys)[i];
    default: return ((const int64_t *)keys)[i];
  }
}

static inline int bit_is_set(const uint8_t *bitmap, uint64_t i) {
  return (bitmap[i >> 3] >> (i & 7)) & 1; /* page/arrow_layout/src/bitmap.rs:9-17 */
}

/* worker_runtime/src/runtime_filter_plan.rs:345-363 (insert_ints) */
uint64_t orc_bloom_insert_keys(const orc_bloom_params *p, uint64_t *bits, const void *keys,
                               int key_width, const uint8_t *validity, uint64_t n) {
  uint64_t rows = 0;
  for (uint64_t i = 0; i < n; ++i) {
    if (validity && !bit_is_set(validity, i)) continue;
    orc_bloom_insert_hash(p, bits, orc_hash_int_key(load_key(keys, key_width, i)));
    ++rows;
  }
  return rows;
}

/* pg/backend_service/src/source.rs:496-532 with a Ready filter of matching generation
 * (runtime_filter/src/shared.rs:350-374). */
uint64_t orc_bloom_probe_keys(const orc_bloom_params *p, const uint64_t *bits, const void *keys,
                              int key_width, const uint8_t *validity, uint64_t n, uint8_t *keep) {
  uint64_t rejected = 0;
  for (uint64_t i = 0; i < n; ++i) {
    int k;
    if (validity && !bit_is_set(validity, i)) {
      k = 0; /* decision_for_null => DefinitelyAbsent */
    } else {
      k = orc_bloom_might_contain_hash(p, bits, orc_hash_int_key(load_key(keys, key_width, i)));
    }
    keep[i] = (uint8_t)k;
    rejected += !k;
  }
  return rejected;
}

/* ---- lifecycle word: runtime_filter/src/shared.rs:7-9 ---- */
#define STATE_BITS 2
#define STATE_MASK 3ull
#define MAX_GENERATION (UINT64_MAX >> STATE_BITS)

/* shared.rs:400-408 */
int orc_lifecycle_pack(uint64_t generation, int state, uint64_t *word) {
  if (generation > MAX_GENERATION) return ORC_LC_GENERATION_EXHAUSTED;
  *word = (generation << STATE_BITS) | (uint64_t)state;
  return ORC_LC_OK;
}

/* shared.rs:411-416 */
void orc_lifecycle_unpack(uint64_t word, uint64_t *generation, int *state) {
  *generation = word >> STATE_BITS;
  *state = (int)(word & STATE_MASK);
}

/* shared.rs:159-198 (single-threaded restatement of the CAS loop) */
int orc_slot_try_acquire_builder(uint64_t *lifecycle, const orc_bloom_params *p, uint64_t *bits,
                                 uint64_t *generation_out) {
  uint64_t gen;
  int state;
  orc_lifecycle_unpack(*lifecycle, &gen, &state);
  if (state == ORC_RF_BUILDING || state == ORC_RF_READY) return ORC_LC_BUSY;
  if (gen == UINT64_MAX || gen + 1 > MAX_GENERATION) return ORC_LC_GENERATION_EXHAUSTED;
  uint64_t w;
  orc_lifecycle_pack(gen + 1, ORC_RF_BUILDING, &w);
  *lifecycle = w;
  orc_bloom_clear(p, bits);
  *generation_out = gen + 1;
  return ORC_LC_OK;
}

/* shared.rs:377-397 (transition_build) */
static int transition_build(uint64_t *lifecycle, uint64_t generation, int next) {
  uint64_t expected, desired;
  orc_lifecycle_pack(generation, ORC_RF_BUILDING, &expected);
  orc_lifecycle_pack(generation, next, &desired);
  if (*lifecycle != expected) return ORC_LC_INVALID_TRANSITION;
  *lifecycle = desired;
  return ORC_LC_OK;
}

int orc_slot_publish_build(uint64_t *lifecycle, uint64_t generation) {
  return transition_build(lifecycle, generation, ORC_RF_READY);
}

int orc_slot_disable_build(uint64_t *lifecycle, uint64_t generation) {
  return transition_build(lifecycle, generation, ORC_RF_DISABLED);
}

/* shared.rs:244-260 */
int orc_slot_retire_ready(uint64_t *lifecycle, uint64_t generation) {
  uint64_t expected, desired;
  if (orc_lifecycle_pack(generation, ORC_RF_READY, &expected)) return ORC_LC_GENERATION_EXHAUSTED;
  orc_lifecycle_pack(generation, ORC_RF_DISABLED, &desired);
  if (*lifecycle != expected) return ORC_LC_INVALID_TRANSITION;
  *lifecycle = desired;
  return ORC_LC_OK;
}

/* shared.rs:350-361 */
int orc_probe_decision_for_hash(const uint64_t *lifecycle, uint64_t generation,
                                const orc_bloom_params *p, const uint64_t *bits, uint64_t hash) {
  uint64_t gen;
  int state;
  orc_lifecycle_unpack(*lifecycle, &gen, &state);
  if (gen != generation || state != ORC_RF_READY) return ORC_PASS_UNFILTERED;
  return orc_bloom_might_contain_hash(p, bits, hash) ? ORC_MAYBE_PRESENT : ORC_DEFINITELY_ABSENT;
}

/* shared.rs:367-374 */
int orc_probe_decision_for_null(const uint64_t *lifecycle, uint64_t generation) {
  uint64_t gen;
  int state;
  orc_lifecycle_unpack(*lifecycle, &gen, &state);
  return (gen == generation && state == ORC_RF_READY) ? ORC_DEFINITELY_ABSENT
                                                      : ORC_PASS_UNFILTERED;
}
