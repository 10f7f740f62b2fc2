"""GPU tests of the Partial -> Final aggregate merge and the Bloom OR-merge inside the library.
With one GPU the "ranks" are shards processed one after the other on the same device (the merge
kernels are identical); with >= 2 GPUs bench.py --gpus N exercises the NCCL path."""
import numpy as np
import pytest
import torch

import pg_fusion_b200 as pg
from oracle import pyorc as O
from pg_fusion_b200 import BloomParams, GenTable
from pg_fusion_b200 import multi_gpu as MG

from . import util as U

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = pg.Context()
    yield c
    c.close()


@pytest.mark.parametrize("shape", ["q6", "q1"])
@pytest.mark.parametrize("world", [2, 4])
def test_partial_states_merge_to_the_single_pass_result(ctx, shape, world):
    rows = 200_000
    table, schema, gpu, orc = ((GenTable.LINEITEM_Q6, U.Q6_SCHEMA, U.gpu_q6, U.oracle_q6) if shape == "q6"
                               else (GenTable.LINEITEM_Q1, U.Q1_SCHEMA, U.gpu_q1, U.oracle_q1))
    stride = 8192
    states = torch.zeros(world * stride, dtype=torch.uint8, device="cuda")
    shards, pages = [], []
    for r in range(world):
        lo, hi = MG.shard_range(rows, r, world)
        s = ctx.gen_scan(table, hi - lo, seed=42, first_row=lo)
        nbytes, stats = gpu(s).run_partial(states.data_ptr() + r * stride, stride)
        assert 0 < nbytes <= stride and stats.rows_in == hi - lo
        shards.append(s)
        pages.append(s.read_pages())
    merged = gpu(shards[0]).merge_partials(states.data_ptr(), stride, world)
    whole = ctx.gen_scan(table, rows, seed=42)
    single = gpu(whole).run()
    want = orc(O.OTable.from_pages(whole.read_pages(), 65536, U.orc_cols(schema)))
    U.assert_agg_equal(merged, want)
    U.assert_agg_equal(merged, single)
    # counts are bit exact
    assert sorted(a[-1] for a in merged.aggs) == sorted(a[-1] for a in want.aggs)
    for s in shards + [whole]:
        s.release()


@pytest.mark.parametrize("shape", ["q6", "q1"])
def test_async_partial_and_bounded_merge(ctx, shape):
    """The one-synchronisation form of the multi-GPU step gives the single-pass result."""
    rows, world, stride = 150_000, 3, 8192
    table, gpu = (GenTable.LINEITEM_Q6, U.gpu_q6) if shape == "q6" else (GenTable.LINEITEM_Q1, U.gpu_q1)
    states = torch.zeros(world * stride, dtype=torch.uint8, device="cuda")
    shards = []
    for r in range(world):
        lo, hi = MG.shard_range(rows, r, world)
        s = ctx.gen_scan(table, hi - lo, seed=42, first_row=lo)
        shards.append(s)
        if r < world - 1:
            gpu(s).run_partial(states.data_ptr() + r * stride, stride)
    last = gpu(shards[-1])
    last.run_partial_async(states.data_ptr() + (world - 1) * stride, stride)   # no synchronisation here
    merged = last.merge_partials_bounded(states.data_ptr(), stride, world)
    lo, hi = MG.shard_range(rows, world - 1, world)
    assert merged.rows_in == hi - lo and merged.kernel_ms > 0
    whole = ctx.gen_scan(table, rows, seed=42)
    U.assert_agg_equal(merged, gpu(whole).run())
    # a stride that cannot carry the groups of a rank is reported, not truncated
    if shape == "q1":
        tiny = 8 * (1 + 19 * 2)   # room for two of the four groups
        small = torch.zeros(tiny, dtype=torch.uint8, device="cuda")
        last.run_partial_async(small.data_ptr(), tiny)
        with pytest.raises(pg.PgfError):
            last.merge_partials_bounded(small.data_ptr(), tiny, 1)
    for s in shards + [whole]:
        s.release()


def test_bloom_or_merge_on_device(ctx):
    p = BloomParams.new(**pg.GUC_DEFAULT_BLOOM)
    keys = np.random.default_rng(2).integers(-2**60, 2**60, 100_000, dtype=np.int64)
    world = 4
    gathered = torch.zeros(world * p.word_count, dtype=torch.int64, device="cuda")
    for r in range(world):
        lo, hi = MG.shard_range(keys.size, r, world)
        rf = ctx.runtime_filter(p)
        rf.try_acquire_builder()
        rf.insert_keys(keys[lo:hi])
        rf.publish_ready()
        gathered[r * p.word_count:(r + 1) * p.word_count] = torch.from_numpy(rf.words().view(np.int64)).cuda()
    torch.cuda.synchronize()
    final = ctx.runtime_filter(p)
    final.try_acquire_builder()
    final.or_device_words(gathered.data_ptr(), world)
    final.publish_ready()
    ob = O.Bloom(O.bloom_params(p.bit_count, p.hash_count, p.seed))
    ob.insert_keys(keys)
    assert (final.words() == ob.words).all()
