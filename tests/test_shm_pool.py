"""Shared-memory runtime-filter pool interop: the worker side of
runtime_filter/src/pool.rs spoken by the library, checked against the pool's binary layout
(pool.rs:152-217), its lifecycle transitions (shared.rs:159-260) and the oracle's Bloom bits.
Follows the reference's pool tests (runtime_filter/src/tests.rs: pool allocate / publish /
probe / release / reuse).  CPU only: the words come from the oracle."""
import ctypes as C
import struct

import numpy as np
import pytest

from oracle import pyorc as O
from pg_fusion_b200 import _lib

MAGIC = 0x5047465552465031
FREE, BUILDING, READY, DISABLED = 0, 1, 2, 3
SLOT_FREE, SLOT_ALLOCATED, SLOT_RETIRING = 0, 1, 2


def params(bits=4096, k=4, seed=42):
    p = _lib.BloomParamsC()
    assert _lib.lib().pgf_bloom_params_new(bits, k, seed, C.byref(p)) == 0
    return p


class Pool:
    def __init__(self, slots, p):
        self.L, self.slots, self.p = _lib.lib(), slots, p
        size, align = C.c_uint64(), C.c_uint64()
        assert self.L.pgf_shm_pool_layout(slots, C.byref(p), C.byref(size), C.byref(align)) == 0
        assert size.value == 56 + 48 * slots + 8 * slots * p.word_count and align.value == 8
        self.buf = np.full(size.value // 8, 0xAB, dtype=np.uint64)   # scratch, not zero: init must clear it
        self.base, self.len = self.buf.ctypes.data_as(C.c_void_p), size.value
        assert self.L.pgf_shm_pool_init(self.base, self.len, slots, C.byref(p)) == 0

    def raw(self):
        return self.buf.view(np.uint8)

    def header(self):
        return struct.unpack_from("<QIIQIIQQQ", self.raw(), 0)

    def slot(self, i):
        state, refs, gen, epoch, scan, col, kt, lifecycle = struct.unpack_from("<IIQQQIIQ", self.raw(), 56 + 48 * i)
        return dict(state=state, refs=refs, generation=gen, session_epoch=epoch, scan_id=scan, output_column=col,
                    key_type=kt, lifecycle=lifecycle)

    def bits(self, i):
        off = (56 + 48 * self.slots) // 8 + i * self.p.word_count
        return self.buf[off:off + self.p.word_count]

    def allocate(self, epoch, scan, col=0, key_type=3):
        t = _lib.RfTarget(epoch, scan, col, key_type)
        slot, gen = C.c_int32(), C.c_uint64()
        rc = self.L.pgf_shm_pool_allocate_build(self.base, self.len, self.slots, C.byref(self.p), C.byref(t), C.byref(slot), C.byref(gen))
        return rc, slot.value, gen.value

    def publish(self, slot, gen, words):
        w = np.ascontiguousarray(words, dtype=np.uint64)
        return self.L.pgf_shm_pool_publish_words(self.base, self.len, self.slots, C.byref(self.p), slot, gen,
                                                 w.ctypes.data_as(C.c_void_p), w.size)

    def release(self, slot):
        return self.L.pgf_shm_pool_release_owner(self.base, self.len, self.slots, C.byref(self.p), slot)


def test_layout_and_header_match_the_reference_structs():
    p = params()
    pool = Pool(3, p)
    magic, version, slots, bit_count, hash_count, _r, seed, word_count, region = pool.header()
    assert (magic, version, slots, bit_count, hash_count, seed, word_count) == (MAGIC, 1, 3, 4096, 4, 42, 64)
    assert region == pool.len
    for i in range(3):
        assert pool.slot(i) == dict(state=SLOT_FREE, refs=0, generation=0, session_epoch=0, scan_id=0, output_column=0, key_type=0, lifecycle=0)
        assert not pool.bits(i).any()
    L = _lib.lib()
    assert L.pgf_shm_pool_attach_check(pool.base, pool.len, 3, C.byref(p)) == 0
    assert L.pgf_shm_pool_attach_check(pool.base, pool.len, 4, C.byref(p)) != 0           # TooSmall / ConfigMismatch
    assert L.pgf_shm_pool_attach_check(pool.base, pool.len, 3, C.byref(params(seed=43))) != 0   # ConfigMismatch
    assert L.pgf_shm_pool_attach_check(C.c_void_p(pool.buf.ctypes.data + 4), pool.len - 4, 3, C.byref(p)) != 0  # Misaligned


def test_allocate_publish_probe_release_and_reuse():
    p = params(bits=1 << 15, k=4, seed=0x7067667573696f6e)
    pool = Pool(2, p)
    keys = np.random.default_rng(5).integers(-2**62, 2**62, 2000, dtype=np.int64)
    ob = O.Bloom(O.bloom_params(p.bit_count, p.hash_count, p.seed))
    ob.insert_keys(keys)

    rc, slot, gen = pool.allocate(epoch=7, scan=11, col=2, key_type=3)
    assert (rc, slot, gen) == (0, 0, 1)
    s = pool.slot(0)
    assert (s["state"], s["refs"], s["generation"], s["session_epoch"], s["scan_id"], s["output_column"], s["key_type"]) == (SLOT_ALLOCATED, 1, 1, 7, 11, 2, 3)
    assert s["lifecycle"] == (1 << 2) | BUILDING
    # a second build takes the next slot; a third finds the pool exhausted (soft miss, not an error)
    assert pool.allocate(7, 12)[:2] == (0, 1)
    assert pool.allocate(7, 13)[:2] == (0, -1)
    # publishing with a stale generation is refused and leaves the slot Building
    assert pool.publish(0, 2, ob.words) != 0
    assert pool.slot(0)["lifecycle"] == (1 << 2) | BUILDING
    assert pool.publish(0, 1, ob.words) == 0
    assert pool.slot(0)["lifecycle"] == (1 << 2) | READY
    assert (pool.bits(0) == ob.words).all() and not pool.bits(1).any()
    # what a backend probe sees: the reference's might_contain over the slot's bits
    probe = O.Bloom(O.bloom_params(p.bit_count, p.hash_count, p.seed))
    probe.words[:] = pool.bits(0)
    assert all(probe.might_contain_u64(int(k) & (2**64 - 1)) for k in keys[:200])
    assert pool.publish(0, 1, ob.words) != 0            # already Ready: not Building any more
    # owner drop with no probe attached: Ready -> Disabled, slot metadata cleared, slot Free again
    assert pool.release(0) == 0
    s = pool.slot(0)
    assert (s["state"], s["refs"], s["generation"], s["session_epoch"], s["scan_id"], s["key_type"]) == (SLOT_FREE, 0, 0, 0, 0, 0)
    assert s["lifecycle"] == (1 << 2) | DISABLED
    # reuse: next generation, bits cleared by the new builder lease
    rc, slot, gen = pool.allocate(8, 21, key_type=2)
    assert (rc, slot, gen) == (0, 0, 2)
    assert pool.slot(0)["lifecycle"] == (2 << 2) | BUILDING and not pool.bits(0).any()
    # a build that fails is disabled, never published
    L = _lib.lib()
    assert L.pgf_shm_pool_disable_build(pool.base, pool.len, 2, C.byref(p), 0, 2) == 0
    assert pool.slot(0)["lifecycle"] == (2 << 2) | DISABLED
    assert pool.publish(0, 2, ob.words) != 0
    assert pool.release(0) == 0 and pool.slot(0)["state"] == SLOT_FREE


def test_release_waits_for_the_last_probe_reference():
    """release_ref (pool.rs:527-556): the slot is only retired by whoever drops the last reference."""
    p = params()
    pool = Pool(1, p)
    rc, slot, gen = pool.allocate(1, 2)
    assert pool.publish(slot, gen, np.ones(p.word_count, dtype=np.uint64)) == 0
    # a backend attached a probe: refs 1 -> 2 (lookup_probes, pool.rs:432-476)
    refs_off = 56 + 4
    struct.pack_into("<I", pool.raw(), refs_off, 2)
    assert pool.release(slot) == 0
    s = pool.slot(0)
    assert (s["state"], s["refs"]) == (SLOT_RETIRING, 1) and s["lifecycle"] == (gen << 2) | READY   # still probe-able
    assert pool.allocate(1, 3)[:2] == (0, -1)   # not reusable yet


# ---- the reference's own pool tests, replayed against this implementation ------------------------
def lookup(pool, epoch, scan):
    out = (_lib.PoolProbe * 8)()
    n = C.c_uint32()
    assert pool.L.pgf_shm_pool_lookup_probes(pool.base, pool.len, pool.slots, C.byref(pool.p), epoch, scan, out, 8, C.byref(n)) == 0
    return [out[i] for i in range(n.value)]


def decide(pool, probe, key=None):
    d = C.c_int32()
    assert pool.L.pgf_shm_pool_probe_decide(pool.base, pool.len, pool.slots, C.byref(pool.p), probe.slot_index, probe.generation,
                                            1 if key is None else 0, 0 if key is None else key, C.byref(d)) == 0
    return d.value


PASS_UNFILTERED, MAYBE_PRESENT, DEFINITELY_ABSENT = 0, 1, 2


def words_for(p, keys):
    b = O.Bloom(O.bloom_params(p.bit_count, p.hash_count, p.seed))
    b.insert_keys(np.array(keys, dtype=np.int64))
    return b.words


def test_pool_publishes_filter_and_probe_rejects_absent_keys():
    # runtime_filter/src/tests.rs:446-478
    p = params(bits=1024, k=3, seed=17)
    pool = Pool(1, p)
    rc, slot, gen = pool.allocate(epoch=11, scan=22, col=3, key_type=3)
    assert (rc, slot) == (0, 0)
    building = lookup(pool, 11, 22)
    assert len(building) == 1 and decide(pool, building[0], 42) == PASS_UNFILTERED      # Building never rejects
    assert pool.L.pgf_shm_pool_release_probe(pool.base, pool.len, 1, C.byref(p), 0) == 0
    assert pool.publish(slot, gen, words_for(p, [42])) == 0
    probes = lookup(pool, 11, 22)
    assert len(probes) == 1 and probes[0].output_column == 3 and probes[0].key_type == 3
    assert decide(pool, probes[0], 42) == MAYBE_PRESENT
    assert decide(pool, probes[0], 100_000) == DEFINITELY_ABSENT
    assert decide(pool, probes[0], None) == DEFINITELY_ABSENT
    assert lookup(pool, 11, 23) == [] and lookup(pool, 12, 22) == []                      # other targets find nothing
    assert pool.slot(0)["refs"] == 2                                                       # owner + one probe


def test_pool_does_not_reuse_storage_until_probes_are_dropped():
    # runtime_filter/src/tests.rs:481-523
    p = params(bits=1024, k=3, seed=17)
    pool = Pool(1, p)
    rc, slot, gen = pool.allocate(epoch=1, scan=2, col=0, key_type=2)
    assert pool.publish(slot, gen, words_for(p, [7])) == 0
    probes = lookup(pool, 1, 2)
    assert len(probes) == 1
    assert pool.release(slot) == 0                          # drop(build)
    assert pool.allocate(3, 4, key_type=2)[:2] == (0, -1)   # allocate while the old probe exists: none
    assert decide(pool, probes[0], 7) == MAYBE_PRESENT      # the old probe still reads consistent bits
    assert pool.L.pgf_shm_pool_release_probe(pool.base, pool.len, 1, C.byref(p), probes[0].slot_index) == 0   # drop(probes)
    rc, slot2, gen2 = pool.allocate(3, 4, key_type=2)
    assert (rc, slot2, gen2) == (0, 0, gen + 1)
    assert decide(pool, probes[0], 7) == PASS_UNFILTERED    # a stale generation never rejects


def test_concurrent_allocation_hands_every_slot_out_once():
    """allocate_build from several threads: the slot-state CAS gives every slot to exactly one builder."""
    import threading
    p = params(bits=4096, k=2, seed=1)
    pool = Pool(16, p)
    got, lock = [], threading.Lock()

    def worker(tid):
        for j in range(8):
            rc, slot, gen = pool.allocate(tid, j)
            assert rc == 0
            if slot >= 0:
                with lock:
                    got.append(slot)
    ts = [threading.Thread(target=worker, args=(t,)) for t in range(8)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert sorted(got) == list(range(16))


def test_pool_protocol_under_thread_sanitizer(tmp_path):
    """tests/cpp/stress_shm_pool.cpp: builder and probe threads hammer a 3-slot pool, csrc/shm_pool.cpp built
    with -fsanitize=thread.  No data race, no false negative, failed builds never reject, every slot reusable
    afterwards -- with lookup_probes serialised against release_owner (the window the reference's unpin leaves
    open), and with allocate_build left concurrent (the window this library closes on the worker side)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "pg_fusion_b200")
    cuda = "/usr/local/cuda"
    obj, exe = str(tmp_path / "shm_pool_tsan.o"), str(tmp_path / "stress_shm_pool")
    san = ["-std=c++17", "-O1", "-g", "-fsanitize=thread", "-I", os.path.join(root, "include")]
    out = subprocess.run(["g++", *san, "-I", os.path.join(cuda, "include"), "-c", os.path.join(lib, "csrc", "shm_pool.cpp"), "-o", obj],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    out = subprocess.run(["g++", *san, "-pthread", "-Wall", "-Wextra", os.path.join(root, "tests", "cpp", "stress_shm_pool.cpp"), obj,
                          "-L", lib, "-lpgf_b200", f"-Wl,-rpath,{lib}", "-L", os.path.join(cuda, "lib64"), "-lcudart",
                          f"-Wl,-rpath,{os.path.join(cuda, 'lib64')}", "-o", exe], capture_output=True, text=True)
    if out.returncode != 0 and "tsan" in out.stderr.lower():
        pytest.skip("thread sanitizer runtime not available")
    assert out.returncode == 0, out.stderr
    for mode in ([], ["release-only"]):
        for _ in range(2):
            run = subprocess.run([exe, "700", *mode], capture_output=True, text=True, timeout=120)
            if "FATAL: ThreadSanitizer" in run.stderr:      # e.g. unsupported address-space layout in this sandbox
                pytest.skip(run.stderr.splitlines()[0])
            assert run.returncode == 0 and "WARNING: ThreadSanitizer" not in run.stderr, run.stdout[-500:] + run.stderr[-3000:]
            assert "false_negatives 0 " in run.stdout and "reusable_slots 3/3" in run.stdout


def test_parameters_and_decisions_match_the_oracle_on_random_inputs():
    """Differential check of the host-side Bloom arithmetic: BloomParams::new / for_expected_items (the
    floating-point sizing formula, bloom.rs:52-79) and the pool's probe decisions (splitmix64 double hashing,
    modulo for arbitrary bit counts) against the oracle, on random parameters, seeds and keys."""
    rng = np.random.default_rng(11)
    L = _lib.lib()
    for _ in range(3000):
        n = int(rng.choice([0, 1, 2, 1000, 10**6, int(rng.integers(1, 2**40)), int(rng.integers(1, 2**63))]))
        fpr = float(rng.choice([0.0, 1.0, -0.1, 1e-12, 0.01, 0.5, 0.999999, float(rng.random()), float("nan"), float("inf")]))
        seed = int(rng.integers(0, 2**63)) * 2 + int(rng.integers(0, 2))
        got = _lib.BloomParamsC()
        rc = L.pgf_bloom_params_for_expected_items(n, fpr, seed, C.byref(got))
        try:
            want = O.bloom_params_for_expected_items(n, fpr, seed)
            assert rc == 0 and (got.bit_count, got.word_count, got.hash_count, got.seed) == \
                (want.bit_count, want.word_count, want.hash_count, want.seed), (n, fpr)
        except O.OracleError as e:
            assert rc == 19 + e.code, (n, fpr, rc, e.code)   # PGF_ERR_BLOOM_* = 20.. in BloomParamError's order (oracle: 1..)
    for _ in range(40):
        bits = int(rng.choice([1, 2, 63, 64, 65, 4096, 4099, int(rng.integers(1, 1 << 16))]))
        k = int(rng.integers(1, 17))
        seed = int(rng.integers(0, 2**63)) * 2 + 1
        p = params(bits, k, seed)
        pool = Pool(2, p)
        keys = rng.integers(-2**63, 2**63 - 1, 200, dtype=np.int64)
        ob = O.Bloom(O.bloom_params(bits, k, seed))
        ob.insert_keys(keys[:100])
        rc, slot, gen = pool.allocate(5, 9)
        assert rc == 0 and pool.publish(slot, gen, ob.words) == 0
        (probe,) = lookup(pool, 5, 9)
        for key in keys:
            want = MAYBE_PRESENT if ob.might_contain_u64(int(key)) else DEFINITELY_ABSENT
            assert decide(pool, probe, int(key)) == want, (bits, k, seed, int(key))
        assert all(decide(pool, probe, int(key)) == MAYBE_PRESENT for key in keys[:100])   # no false negatives
