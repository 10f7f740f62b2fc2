"""World-size-2 gloo tests (CPU) of the N > 1 host logic: page sharding, Bloom OR-merge after an
all-gather and Partial -> Final aggregate merging.  The per-rank compute is done by the oracle
here (no GPU); the same merges run inside libpgf_b200 on the GPUs (tests/test_gpu_multi.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pyorc as O
from pg_fusion_b200 import multi_gpu as MG

from . import util as U


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # --- Bloom: every rank inserts its shard of the build keys, words are all-gathered and OR-ed
        keys = np.random.default_rng(3).integers(0, 10**9, 40_000, dtype=np.int64)
        lo, hi = MG.shard_range(keys.size, rank, world)
        p = O.bloom_params(1 << 18, 4, 17)
        mine = O.Bloom(p)
        mine.insert_keys(keys[lo:hi])
        gathered = MG.all_gather_bytes(torch.from_numpy(mine.words.view(np.uint8).copy()), world)
        merged = MG.or_merge_words(gathered.numpy(), world)
        whole = O.Bloom(p)
        whole.insert_keys(keys)
        ok_bloom = bool((merged == whole.words).all())
        # --- Q1 shape: pages sharded by contiguous page range, partial states merged in rank order
        li = U.lineitem(20_000, 5)
        pages = U.q1_pages(li)
        plo, phi = MG.shard_range(pages.shape[0], rank, world)
        part = U.oracle_q1(O.OTable.from_pages(pages[plo:phi], 65536, U.orc_cols(U.Q1_SCHEMA)))
        local = {k: ((a[0], a[7]), ) for k, a in part.by_key().items()}
        objs = [None] * world
        dist.all_gather_object(objs, {k: (v[0][0], v[0][1]) for k, v in local.items()})
        final = MG.merge_partial_sums(objs)
        single = U.oracle_q1(O.OTable.from_pages(pages, 65536, U.orc_cols(U.Q1_SCHEMA))).by_key()
        ok_counts = all(final[k][1] == single[k][7] for k in single) and set(final) == set(single)
        ok_sums = all(abs(final[k][0] - single[k][0]) <= 1e-12 * abs(single[k][0]) for k in single)
        # every rank computes the same merged result (fixed merge order)
        t = torch.tensor([final[k][0] for k in sorted(final)], dtype=torch.float64)
        ts = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(ts, t)
        ok_same = all(torch.equal(ts[0], x) for x in ts)
        q.put((rank, ok_bloom, ok_counts, ok_sums, ok_same))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 37167, 10**9 + 7):
        for world in (1, 2, 3, 8):
            spans = [MG.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    with pytest.raises(ValueError):
        MG.shard_range(10, 2, 2)


def test_two_rank_gloo_bloom_or_and_partial_final_merge():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_bloom, ok_counts, ok_sums, ok_same in results:
        assert ok_bloom, f"rank {rank}: OR of shard bit arrays != bit array of the union"
        assert ok_counts and ok_sums and ok_same, f"rank {rank}: partial/final merge mismatch"
