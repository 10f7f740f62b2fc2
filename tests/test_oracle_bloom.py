"""Pins the Bloom oracle (oracle/orc_bloom.c) against the reference's own known-answer
tests: runtime_filter/src/tests.rs (line numbers cited per test)."""
import hashlib

import numpy as np
import pytest

from oracle import pyorc as O

M64 = 2**64 - 1


def test_splitmix64_known_values():
    # SplitMix64 (Steele/Lea/Flood) first outputs for state 0: the generator adds the gamma
    # then mixes, which is exactly bloom.rs:293-299 applied to 0, gamma, 2*gamma ...
    L = O.lib()
    assert L.orc_splitmix64(0) == 0xE220A8397B1DCDAF
    assert L.orc_splitmix64(0x9E3779B97F4A7C15) == 0x6E789E6AA1B965F4
    assert L.orc_splitmix64((2 * 0x9E3779B97F4A7C15) & M64) == 0x06C45D188009454F


def test_hash_int_key_is_sign_extending_identity():
    L = O.lib()  # runtime_filter/src/lib.rs:31-34
    assert L.orc_hash_int_key(42) == 42
    assert L.orc_hash_int_key(-1) == M64
    assert L.orc_hash_int_key(-(2**63)) == 2**63


def test_no_false_negatives_for_inserted_keys():
    # tests.rs:48-61 ; for_expected_items(1000, 0.01, 0xA5A5) => 9586 bits / 7 hashes
    p = O.bloom_params_for_expected_items(1000, 0.01, 0xA5A5)
    assert (p.bit_count, p.hash_count, p.word_count) == (9586, 7, 150)
    b = O.Bloom(p)
    for k in range(1000):
        b.insert_u64(k)
    assert all(b.might_contain_u64(k) for k in range(1000))


def test_builder_lease_publishes_ready_filter():
    # tests.rs:63-83
    s = O.Slot(O.bloom_params(512, 4, 42))
    rc, gen = s.try_acquire_builder()
    assert (rc, gen) == (O.lib().orc_lifecycle_pack and 0, 1)
    s.bloom.insert_u64(10)
    assert s.publish(gen) == 0
    assert s.snapshot() == (1, O.RF_READY)
    assert s.decision_for_u64(1, 10) == O.MAYBE_PRESENT
    assert s.decision_for_u64(1, 99) == O.DEFINITELY_ABSENT


def test_free_building_disabled_and_stale_generations_never_reject():
    # tests.rs:85-127
    s = O.Slot(O.bloom_params(256, 3, 0))
    assert s.decision_for_u64(1, 99) == O.PASS_UNFILTERED
    rc, gen = s.try_acquire_builder()
    s.bloom.insert_u64(1)
    assert s.decision_for_u64(gen, 99) == O.PASS_UNFILTERED
    assert s.disable(gen) == 0  # dropped builder disables its generation (shared.rs:319-325)
    assert s.snapshot() == (1, O.RF_DISABLED)
    assert s.decision_for_u64(1, 99) == O.PASS_UNFILTERED
    rc, gen = s.try_acquire_builder()
    assert (rc, gen) == (0, 2)
    s.bloom.insert_u64(2)
    assert s.publish(gen) == 0
    assert s.decision_for_u64(1, 2) == O.PASS_UNFILTERED
    assert s.decision_for_u64(2, 2) == O.MAYBE_PRESENT


def test_second_builder_rejected_and_ready_not_reused():
    # tests.rs:129-173
    s = O.Slot(O.bloom_params(256, 3, 0))
    rc, gen = s.try_acquire_builder()
    assert s.try_acquire_builder()[0] == 2  # Busy while Building
    s.bloom.insert_u64(7)
    assert s.publish(gen) == 0
    assert s.try_acquire_builder()[0] == 2  # Busy while Ready
    assert s.decision_for_u64(gen, 7) == O.MAYBE_PRESENT


def test_quiescent_retire_allows_reuse_and_stale_retire_fails():
    # tests.rs:175-245
    s = O.Slot(O.bloom_params(256, 3, 0))
    rc, gen = s.try_acquire_builder()
    s.bloom.insert_u64(7)
    s.publish(gen)
    assert s.retire_ready(0) == 3  # InvalidTransition
    assert s.retire_ready(gen) == 0
    assert s.decision_for_u64(gen, 7) == O.PASS_UNFILTERED
    rc, gen2 = s.try_acquire_builder()
    assert (rc, gen2) == (0, 2)
    assert s.decision_for_u64(gen, 7) == O.PASS_UNFILTERED
    s.bloom.insert_u64(9)
    s.publish(gen2)
    assert s.decision_for_u64(gen2, 9) == O.MAYBE_PRESENT
    assert s.retire_ready(1) == 3


def test_stale_builder_transition_cannot_overwrite_newer_generation():
    # tests.rs:292-321
    import ctypes as C
    s = O.Slot(O.bloom_params(256, 3, 0))
    rc, gen = s.try_acquire_builder()
    w = C.c_uint64()
    O.lib().orc_lifecycle_pack(2, O.RF_READY, C.byref(w))
    s.word.value = w.value
    assert s.disable(gen) == 3
    assert s.snapshot() == (2, O.RF_READY)


def test_max_generation_free_slot_is_exhausted():
    # tests.rs:323-340
    import ctypes as C
    s = O.Slot(O.bloom_params(64, 2, 0))
    w = C.c_uint64()
    assert O.lib().orc_lifecycle_pack(M64 >> 2, O.RF_FREE, C.byref(w)) == 0
    s.word.value = w.value
    assert s.try_acquire_builder()[0] == 1  # GenerationExhausted
    assert O.lib().orc_lifecycle_pack((M64 >> 2) + 1, O.RF_FREE, C.byref(w)) == 1


def test_tiny_filter_boundaries():
    # tests.rs:342-352
    p = O.bloom_params(1, 8, M64)
    assert p.word_count == 1
    b = O.Bloom(p)
    b.clear()
    b.insert_u64(123)
    assert b.might_contain_u64(123)
    assert int(b.words[0]) == 1


def test_parameter_validation_rejects_invalid_inputs():
    # tests.rs:354-372
    for args, code in (((0, 1, 0), 1), ((1, 0, 0), 2)):
        with pytest.raises(O.OracleError) as e:
            O.bloom_params(*args)
        assert e.value.code == code
    with pytest.raises(O.OracleError) as e:
        O.bloom_params_for_expected_items(0, 0.01, 0)
    assert e.value.code == 3
    with pytest.raises(O.OracleError) as e:
        O.bloom_params_for_expected_items(10, 1.0, 0)
    assert e.value.code == 4


def test_lifecycle_snapshot_roundtrip():
    # tests.rs:374-384
    import ctypes as C
    w = C.c_uint64()
    assert O.lib().orc_lifecycle_pack(123, O.RF_DISABLED, C.byref(w)) == 0
    assert w.value == (123 << 2) | 3
    g, s = C.c_uint64(), C.c_int()
    O.lib().orc_lifecycle_unpack(w.value, C.byref(g), C.byref(s))
    assert (g.value, s.value) == (123, O.RF_DISABLED)


def test_attach_requires_enough_words():
    # tests.rs:386-398
    p = O.bloom_params(129, 3, 0)
    assert p.word_count == 3
    with pytest.raises(O.OracleError) as e:
        O.Bloom(p, np.zeros(2, dtype=np.uint64))
    assert e.value.code == 6


def test_statistical_false_positive_rate():
    # tests.rs:400-422
    p = O.bloom_params_for_expected_items(1000, 0.01, 0xB10F)
    b = O.Bloom(p)
    b.insert_keys(np.arange(1000, dtype=np.int64))
    keep, rejected = b.probe_keys(np.arange(10_000, 20_000, dtype=np.int64))
    fp = int(keep.sum())
    assert fp < 400
    assert fp + rejected == 10_000
    assert fp == 100  # value observed by SURVEY 8c's independent restatement


def test_pool_publishes_filter_and_probe_rejects_absent_keys():
    # tests.rs:445-478 (pool fixture params new(1024,3,17), tests.rs:27-28)
    s = O.Slot(O.bloom_params(1024, 3, 17))
    rc, gen = s.try_acquire_builder()
    s.bloom.insert_u64(O.lib().orc_hash_int_key(42))
    s.publish(gen)
    assert s.decision_for_u64(gen, 42) == O.MAYBE_PRESENT
    assert s.decision_for_u64(gen, 100_000) == O.DEFINITELY_ABSENT
    assert s.decision_for_null(gen) == O.DEFINITELY_ABSENT
    assert [s.bloom.bit_index(42, i) for i in range(3)] == [60, 723, 362]


def test_bit_positions_derived_vectors():
    # vectors derived independently from bloom.rs:244-255,293-299
    b = O.Bloom(O.bloom_params(512, 4, 42))
    assert [b.bit_index(10, i) for i in range(4)] == [1, 306, 99, 404]
    b.insert_u64(10)
    assert [int(w) for w in b.words] == [0x2, 0x800000000, 0, 0, 0x4000000000000, 0, 0x100000, 0]
    d = O.Bloom(O.bloom_params(1 << 20, 4, 0x7067667573696F6E))
    assert [d.bit_index(1, i) for i in range(4)] == [179616, 13537, 896034, 729955]
    assert [d.bit_index(-1, i) for i in range(4)] == [831129, 632222, 433315, 234408]


def test_guc_default_bit_array_digest():
    # GUC defaults: pg/extension/src/guc.rs:41-46 (1 048 576 bits, 4 hashes, seed "pgfusion")
    p = O.bloom_params(1 << 20, 4, 0x7067667573696F6E)
    b = O.Bloom(p)
    assert p.word_count == 16384
    n = b.insert_keys(np.arange(1, 1001, dtype=np.int64))
    assert n == 1000
    assert int(sum(bin(int(w)).count("1") for w in b.words)) == 3994
    digest = hashlib.sha256(b.words.astype("<u8").tobytes()).hexdigest()
    assert digest == "748e72253a2859a687a1452a75ee832856704e4ea717e4cd7305e78e177f8a3c"


def test_key_widening_and_nulls():
    # Int16/Int32 keys sign-extend to i64 (runtime_filter_plan.rs:244,256,268); NULL keys are
    # not inserted (:345-363) and probe as DefinitelyAbsent (shared.rs:367-374)
    p = O.bloom_params(4096, 4, 7)
    k16 = np.array([-3, 5, 7, -32768], dtype=np.int16)
    b16, b32, b64 = O.Bloom(p), O.Bloom(p), O.Bloom(p)
    b16.insert_keys(k16)
    b32.insert_keys(k16.astype(np.int32))
    b64.insert_keys(k16.astype(np.int64))
    assert (b16.words == b64.words).all() and (b32.words == b64.words).all()
    validity = np.array([0b0101], dtype=np.uint8)  # rows 0 and 2 valid
    bn = O.Bloom(p)
    assert bn.insert_keys(k16, validity) == 2
    keep, rejected = bn.probe_keys(k16, validity)
    assert keep.tolist()[0] == 1 and keep.tolist()[2] == 1
    assert keep.tolist()[1] == 0 and keep.tolist()[3] == 0  # NULL => absent
    assert rejected == 2
