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


def _partition_worker(rank, world, port, q):
    """The host logic of the hash-partitioned Q3 plan (tpch.gpu_q3_partitioned) on two gloo ranks, the oracle doing
    the arithmetic: rows are owned by pgf_partition_of_key (the library's routing function, host callable), every rank
    routes the rows of its page shard to their owners, joins and groups what it received, and the ranks' top rows
    are merged with pack_topk / merge_topk.  The result must be the single-process oracle's."""
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pg_fusion_b200 import _lib
        from pg_fusion_b200 import tpch as T
        owner = lambda k: _lib.lib().pgf_partition_of_key(int(k), world)
        pages, tables = U.q3_host_tables(300, 2000, 12_000, seed=3, rows_per_page=400)
        want, _ = U.oracle_q3(*tables)
        cust_t, ord_t, li_t = tables
        # this rank's page shards, decoded to rows (customer is broadcast: every rank sees all of it)
        def shard(t, pg):
            lo, hi = MG.shard_range(pg.shape[0], rank, world)
            return O.OTable.from_pages(pg[lo:hi], 65536, t)
        my_orders = shard(U.orc_cols(U.ORDERS_SCHEMA), pages[1])
        my_li = shard(U.orc_cols(U.LINEITEM_Q3_SCHEMA), pages[2])
        building = {int(k) for k, s in zip(cust_t.column(0)[0], cust_t.column(1)) if s == b"BUILDING"}
        # orders: filter, join with customer, route by owner of o_orderkey
        ok, oc, od, op = my_orders.column(0)[0], my_orders.column(1)[0], my_orders.column(2), my_orders.column(3)[0]
        out_o = [[] for _ in range(world)]
        for k, c, d, p in zip(ok, oc, od, op):
            if d < U.Q3_DATE and int(c) in building:
                out_o[owner(k)].append((int(k), bytes(d), int(p)))
        recv = [None] * world
        dist.all_gather_object(recv, out_o)                       # gloo stand-in for the NCCL all-to-all
        mine_o = [row for r in range(world) for row in recv[r][rank]]
        assert all(owner(k) == rank for k, _, _ in mine_o)
        table = {}
        for k, d, p in mine_o:
            table.setdefault(k, []).append((d, p))
        # lineitem: filter, route by owner of l_orderkey (the runtime filter would drop most rows before this point)
        lk, lp, ld, ls = my_li.column(0)[0], my_li.column(1)[0], my_li.column(2)[0], my_li.column(3)
        out_l = [[] for _ in range(world)]
        for k, p, d, s in zip(lk, lp, ld, ls):
            if s > U.Q3_DATE:
                out_l[owner(k)].append((int(k), float(p), float(d)))
        recv = [None] * world
        dist.all_gather_object(recv, out_l)
        groups = {}
        for r in range(world):                                   # rank order, then row order: a fixed summation order
            for k, p, d in recv[r][rank]:
                for od_, op_ in table.get(k, ()):
                    key = (k, od_, op_)
                    groups[key] = groups.get(key, 0.0) + p * (1.0 - d)
        rows = sorted(((k[0], v, k[1], k[2]) for k, v in groups.items()), key=lambda r: (-r[1], r[2]))
        bufs = [None] * world
        dist.all_gather_object(bufs, T.pack_topk(rows, 10))
        top = T.merge_topk(bufs, 10)
        want10 = U.top10(want)
        ok_top = [(r[0], r[2], r[3]) for r in top] == [(r[0], r[2], r[3]) for r in want10] and \
            all(abs(a[1] - b[1]) <= 1e-12 * abs(b[1]) for a, b in zip(top, want10))
        ref = want.by_key()
        ok_groups = all(owner(k[0]) == rank and abs(v - ref[k][0]) <= 1e-12 * abs(ref[k][0]) for k, v in groups.items())
        counts = [None] * world
        dist.all_gather_object(counts, len(groups))
        q.put((rank, ok_top, ok_groups, sum(counts) == len(ref)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_hash_partitioned_q3_plan():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_partition_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_top, ok_groups, ok_cover in results:
        assert ok_top, f"rank {rank}: merged top-10 differs from the oracle"
        assert ok_groups, f"rank {rank}: a group is not owned by this rank or its sum differs"
        assert ok_cover, "the ranks' groups do not add up to the oracle's group count"
