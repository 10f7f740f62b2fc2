"""GPU parity tests of the runtime Bloom filter kernels, through the C ABI, against the
oracle (oracle/orc_bloom.c) and the reference's own known-answer tests
(runtime_filter/src/tests.rs).  Bar: bit arrays and probe decisions bit-exact."""
import hashlib

import numpy as np
import pytest

import pg_fusion_b200 as pg
from oracle import pyorc as O
from pg_fusion_b200 import BloomParams, ColumnSpec, ProbeDecision, RuntimeFilterState, TypeTag
from pg_fusion_b200 import arrow_layout as AL

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = pg.Context()
    yield c
    c.close()


def oracle_bloom(p: BloomParams) -> O.Bloom:
    return O.Bloom(O.bloom_params(p.bit_count, p.hash_count, p.seed))


def built(ctx, params, keys, validity=None):
    rf = ctx.runtime_filter(params)
    rf.try_acquire_builder()
    n = rf.insert_keys(keys, validity)
    rf.publish_ready()
    return rf, n


def test_builder_lease_publishes_ready_filter(ctx):
    # runtime_filter/src/tests.rs:63-83
    rf = ctx.runtime_filter(BloomParams.new(512, 4, 42))
    assert rf.try_acquire_builder() == 1
    rf.insert_u64(10)
    rf.publish_ready()
    assert rf.snapshot() == (1, RuntimeFilterState.Ready)
    assert rf.decision_for_u64(10) == ProbeDecision.MaybePresent
    assert rf.decision_for_u64(99) == ProbeDecision.DefinitelyAbsent
    assert [int(w) for w in rf.words()] == [0x2, 0x800000000, 0, 0, 0x4000000000000, 0, 0x100000, 0]


def test_lifecycle_matches_reference_state_machine(ctx):
    # tests.rs:85-127,129-173,175-215,247-290
    rf = ctx.runtime_filter(BloomParams.new(256, 3, 0))
    assert rf.decision_for_u64(99, generation=1) == ProbeDecision.PassUnfiltered  # Free
    g = rf.try_acquire_builder()
    rf.insert_u64(1)
    assert rf.decision_for_u64(99, generation=g) == ProbeDecision.PassUnfiltered  # Building
    with pytest.raises(pg.PgfError) as e:
        rf.try_acquire_builder()
    assert e.value.name == "LIFECYCLE_BUSY"
    rf.disable()
    assert rf.snapshot() == (1, RuntimeFilterState.Disabled)
    assert rf.decision_for_u64(99, generation=1) == ProbeDecision.PassUnfiltered
    assert rf.try_acquire_builder() == 2
    assert int(rf.words().sum()) == 0  # acquire clears the payload
    rf.insert_u64(2)
    rf.publish_ready()
    assert rf.decision_for_u64(2, generation=1) == ProbeDecision.PassUnfiltered  # stale generation
    assert rf.decision_for_u64(2) == ProbeDecision.MaybePresent
    assert rf.decision_for_null() == ProbeDecision.DefinitelyAbsent
    with pytest.raises(pg.PgfError) as e:
        rf.try_acquire_builder()  # Ready slots are not reused without retire
    assert e.value.name == "LIFECYCLE_BUSY"
    with pytest.raises(pg.PgfError) as e:
        rf.disable()
    assert e.value.name == "LIFECYCLE_INVALID_TRANSITION"
    rf.retire_ready_after_quiescence()
    assert rf.decision_for_u64(2) == ProbeDecision.PassUnfiltered
    assert rf.try_acquire_builder() == 3
    with pytest.raises(pg.PgfError):
        ctx.runtime_filter(BloomParams.new(64, 2, 0)).insert_u64(5)  # insert requires Building


def test_no_false_negatives_and_fp_rate(ctx):
    # tests.rs:48-61,400-422
    p = BloomParams.for_expected_items(1000, 0.01, 0xB10F)
    assert (p.bit_count, p.hash_count) == (9586, 7)
    keys = np.arange(1000, dtype=np.int64)
    rf, n = built(ctx, p, keys)
    assert n == 1000
    d, st = rf.probe_keys(keys)
    assert (d == ProbeDecision.MaybePresent).all() and st.rejected_rows == 0
    d, st = rf.probe_keys(np.arange(10_000, 20_000, dtype=np.int64))
    fp = int((d == ProbeDecision.MaybePresent).sum())
    assert fp < 400 and fp == 100 and st.rejected_rows == 10_000 - fp and st.probe_rows == 10_000


def test_tiny_filter_and_attach_words(ctx):
    # tests.rs:342-352,386-398
    rf, _ = built(ctx, BloomParams.new(1, 8, 2**64 - 1), np.array([123], dtype=np.int64))
    assert rf.words().tolist() == [1]
    assert rf.decision_for_u64(123) == ProbeDecision.MaybePresent
    rf3 = ctx.runtime_filter(BloomParams.new(129, 3, 0))
    assert rf3.params.word_count == 3
    from pg_fusion_b200 import _lib
    import ctypes as C
    two = np.zeros(2, dtype=np.uint64)
    rc = _lib.lib().pgf_bloom_read_words(ctx.h, rf3.handle, two.ctypes.data_as(C.c_void_p), 2)
    assert rc == 25  # InsufficientWords


def test_guc_default_bit_array_digest(ctx):
    p = BloomParams.new(**pg.GUC_DEFAULT_BLOOM)
    rf, n = built(ctx, p, np.arange(1, 1001, dtype=np.int64))
    w = rf.words()
    assert n == 1000 and w.size == 16384
    assert hashlib.sha256(w.astype("<u8").tobytes()).hexdigest() == \
        "748e72253a2859a687a1452a75ee832856704e4ea717e4cd7305e78e177f8a3c"


@pytest.mark.parametrize("bits,k,seed", [(1 << 20, 4, 0x7067667573696F6E), (9586, 7, 0xA5A5), (1000003, 5, 1),
                                         (129, 3, 0), (1 << 24, 2, 99), (2**33 + 7, 3, 5), (64, 1, 2**64 - 1)])
@pytest.mark.parametrize("dtype", [np.int64, np.int32, np.int16])
def test_build_and_probe_bit_exact_vs_oracle(ctx, bits, k, seed, dtype):
    if bits > 1 << 28:
        pytest.skip("bit array too large for the oracle to be practical") if bits > 2**34 else None
    r = np.random.default_rng(bits % 1000 + k)
    info = np.iinfo(dtype)
    n = 20_000
    keys = r.integers(info.min, info.max, n, dtype=dtype, endpoint=True)
    keys[:4] = [info.min, info.max, 0, -1]
    validity = r.random(n) > 0.1
    p = BloomParams.new(bits, k, seed)
    ob = oracle_bloom(p)
    vb = np.packbits(validity, bitorder="little")
    want_n = ob.insert_keys(keys, vb)
    rf, got_n = built(ctx, p, keys, validity)
    assert got_n == want_n
    assert (rf.words() == ob.words).all()
    probe = np.concatenate([keys[:5000], r.integers(info.min, info.max, 5000, dtype=dtype, endpoint=True)])
    pvalid = r.random(probe.size) > 0.05
    keep, rejected = ob.probe_keys(probe, np.packbits(pvalid, bitorder="little"))
    d, st = rf.probe_keys(probe, pvalid)
    assert ((d == ProbeDecision.MaybePresent) == (keep == 1)).all()
    assert ((d == ProbeDecision.DefinitelyAbsent) == (keep == 0)).all()
    assert st.rejected_rows == rejected and st.pass_unfiltered == 0


def test_scan_build_and_probe_over_pages(ctx):
    """Keys arrive as page-backed Int64/Int32 columns (8056 rows per Int64 page)."""
    r = np.random.default_rng(5)
    n = 30_000
    keys = r.integers(-2**62, 2**62, n, dtype=np.int64)
    valid = r.random(n) > 0.2
    schema = [ColumnSpec(TypeTag.Int64, True), ColumnSpec(TypeTag.Int32, False)]
    k32 = r.integers(-2**31, 2**31 - 1, n, dtype=np.int32)
    pages = AL.encode_pages(schema, [(keys, valid), (k32, None)])
    scan = ctx.declare_scan(schema)
    scan.push_pages(pages)
    scan.finish()
    assert scan.info().rows == n
    p = BloomParams.new(**pg.GUC_DEFAULT_BLOOM)
    for col, kk, vv in ((0, keys, valid), (1, k32, None)):
        ob = oracle_bloom(p)
        want = ob.insert_keys(kk, None if vv is None else np.packbits(vv, bitorder="little"))
        rf = ctx.runtime_filter(p)
        rf.try_acquire_builder()
        assert rf.insert_scan(scan, col) == want
        rf.publish_ready()
        assert (rf.words() == ob.words).all()
        d, st = rf.probe_scan(scan, col)
        keep, rejected = ob.probe_keys(kk, None if vv is None else np.packbits(vv, bitorder="little"))
        assert ((d == ProbeDecision.MaybePresent) == (keep == 1)).all()
        assert st.rejected_rows == rejected and st.probe_rows == n
    scan.release()


def test_or_merge_of_shards_equals_single_build(ctx):
    """Multi-GPU merge rule: OR of per-shard bit arrays == bit array of the union."""
    r = np.random.default_rng(9)
    keys = r.integers(0, 10**9, 50_000, dtype=np.int64)
    p = BloomParams.new(1 << 18, 4, 17)
    whole, _ = built(ctx, p, keys)
    a, _ = built(ctx, p, keys[:20_000])
    b, _ = built(ctx, p, keys[20_000:])
    merged = ctx.runtime_filter(p)
    merged.try_acquire_builder()
    merged.or_words(a.words())
    merged.or_words(b.words())
    merged.publish_ready()
    assert (merged.words() == whole.words()).all()


def test_empty_inputs(ctx):
    rf, n = built(ctx, BloomParams.new(1024, 3, 17), np.zeros(0, dtype=np.int64))
    assert n == 0 and int(rf.words().sum()) == 0
    d, st = rf.probe_keys(np.zeros(0, dtype=np.int64))
    assert d.size == 0 and st.probe_rows == 0


def test_gpu_built_filter_is_published_into_the_shared_memory_pool(ctx):
    """RuntimeFilterBuildExec on the GPU, probes by unchanged backends: the words
    built in HBM land bit-exact in the pool slot and the slot goes Building -> Ready."""
    import ctypes as C
    import struct
    from pg_fusion_b200 import _lib
    p = BloomParams.new(**pg.GUC_DEFAULT_BLOOM)
    pc = _lib.BloomParamsC(p.bit_count, p.word_count, p.hash_count, p.seed)
    L, slots = _lib.lib(), 4
    size = C.c_uint64()
    assert L.pgf_shm_pool_layout(slots, C.byref(pc), C.byref(size), None) == 0
    buf = np.zeros(size.value // 8, dtype=np.uint64)
    base = buf.ctypes.data_as(C.c_void_p)
    assert L.pgf_shm_pool_init(base, size.value, slots, C.byref(pc)) == 0
    target = _lib.RfTarget(3, 99, 1, 3)
    slot, gen = C.c_int32(), C.c_uint64()
    assert L.pgf_shm_pool_allocate_build(base, size.value, slots, C.byref(pc), C.byref(target), C.byref(slot), C.byref(gen)) == 0
    keys = np.random.default_rng(9).integers(-2**62, 2**62, 300_000, dtype=np.int64)
    rf, n = built(ctx, p, keys)
    rf.publish_to_pool(buf.ctypes.data, size.value, slots, slot.value, gen.value)
    ob = oracle_bloom(p)
    ob.insert_keys(keys)
    off = (56 + 48 * slots) // 8 + slot.value * p.word_count
    assert (buf[off:off + p.word_count] == ob.words).all()
    lifecycle = struct.unpack_from("<Q", buf.view(np.uint8), 56 + 48 * slot.value + 40)[0]
    assert lifecycle == (gen.value << 2) | int(RuntimeFilterState.Ready)
    with pytest.raises(pg.PgfError):   # a second publish of the same generation is refused
        rf.publish_to_pool(buf.ctypes.data, size.value, slots, slot.value, gen.value)


def test_runtime_filter_metric_counters():
    """The RuntimeFilter* counters of the reference's registry (runtime_metrics/src/lib.rs:125-131), kept per context:
    allocated / ready / pool exhausted / build rows (worker_runtime/src/runtime_filter_plan.rs:89-92,272,283) and
    probe rows / rejected / pass-unfiltered (pg/backend_service/src/source.rs:474-493) -- from the stand-alone Bloom
    calls and from the probes fused into pgf_pipeline_run."""
    with pg.Context() as c:
        assert set(c.runtime_filter_metrics().values()) == {0}
        r = np.random.default_rng(21)
        n = 40_000
        build_keys = np.arange(0, 2000, dtype=np.int64)
        probe_keys = r.integers(0, 20_000, n).astype(np.int64)
        vals = r.random(n)
        p = BloomParams.new(1 << 16, 4, 5)
        rf = c.runtime_filter(p)
        rf.try_acquire_builder()
        assert c.runtime_filter_metrics()["allocated_total"] == 1
        assert rf.insert_keys(build_keys) == 2000
        # a probe before Ready passes every row unfiltered (shared.rs:350-361)
        d, st = rf.probe_keys(probe_keys)
        assert st.pass_unfiltered == n and st.rejected_rows == 0
        rf.publish_ready()
        d, st = rf.probe_keys(probe_keys)
        m = c.runtime_filter_metrics()
        assert (m["ready_total"], m["build_rows_total"]) == (1, 2000)
        assert m["probe_rows_total"] == 2 * n and m["probe_pass_unfiltered_total"] == n and m["probe_rows_rejected_total"] == st.rejected_rows > 0
        # the same filter fused into a scan pipeline: every scanned row is probed, the rejected ones never reach the predicate
        schema = [ColumnSpec(TypeTag.Int64), ColumnSpec(TypeTag.Float64)]
        scan = c.declare_scan(schema)
        scan.push_pages(AL.encode_pages(schema, [(probe_keys, None), (vals, None)]))
        scan.finish()
        res = scan.pipeline().bloom_probe(rf, 0).filter(1, pg.Cmp.GE, 0.0).aggregate([], [(pg.AggFunc.COUNT_STAR, None)]).run()
        m2 = c.runtime_filter_metrics()
        assert res.rows_in == n and res.rows_in - res.rows_bloom == st.rejected_rows
        assert m2["probe_rows_total"] - m["probe_rows_total"] == n
        assert m2["probe_rows_rejected_total"] - m["probe_rows_rejected_total"] == st.rejected_rows
        # a retired filter cannot be consulted: the fused probe is dropped and its rows pass unfiltered
        rf.retire_ready_after_quiescence()
        res = scan.pipeline().bloom_probe(rf, 0).aggregate([], [(pg.AggFunc.COUNT_STAR, None)]).run()
        m3 = c.runtime_filter_metrics()
        assert res.aggs[0][0] == n and m3["probe_pass_unfiltered_total"] - m2["probe_pass_unfiltered_total"] == n
        assert m3["probe_rows_total"] == m2["probe_rows_total"]
        c.note_pool_exhausted()     # what the planner hook reports when allocate_build finds no free slot
        assert c.runtime_filter_metrics()["pool_exhausted_total"] == 1
        scan.release()
