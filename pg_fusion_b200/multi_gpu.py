"""Host-side plumbing for one-process-per-GPU runs (torch.distributed: NCCL on GPUs, gloo in
the CPU tests).  No compute happens here: shards are page ranges, merges run in the library.

Sharding mirrors the reference's scan-side parallelism, where dynamic PostgreSQL workers scan
disjoint CTID block ranges and each produces its own page stream (ai/architecture.md:119-133):
here every rank owns a contiguous range of pages of each scan.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) share of n units (pages or rows) for `rank` of `world`."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return n * rank // world, n * (rank + 1) // world


def all_gather_bytes(local, world: int):
    """All-gather a fixed-size uint8 tensor: returns a [world * n] tensor on the same device."""
    import torch
    import torch.distributed as dist
    out = torch.empty(world * local.numel(), dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())
    return out


def or_merge_words(gathered: np.ndarray, world: int) -> np.ndarray:
    """Bitwise OR of `world` Bloom word arrays laid out back to back (idempotent, order free)."""
    w = np.ascontiguousarray(gathered).view(np.uint64).reshape(world, -1)
    return np.bitwise_or.reduce(w, axis=0)


def merge_partial_sums(parts: Sequence[dict]) -> dict:
    """Reference-order merge of per-rank partial aggregate states {key: (sum, count)} in rank
    order (AggregateExec Partial -> Final): exact for integers, fixed order for floats."""
    out: dict = {}
    for p in parts:
        for k, (s, c) in p.items():
            if k in out:
                out[k] = (out[k][0] + s, out[k][1] + c)
            else:
                out[k] = (s, c)
    return out


# ---- collectives over torch.distributed (NCCL on GPUs); the compute stays in the library ----
def all_gather_counts(n: int, world: int, device) -> List[int]:
    import torch
    import torch.distributed as dist
    t = torch.tensor([int(n)], dtype=torch.int64, device=device)
    out = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, t)
    return [int(x) for x in out.tolist()]


def broadcast_join_table(ctx, handle: int, world: int, device) -> int:
    """Broadcast join: every rank exports the rows of the table it built from its own
    pages, the fragments are all-gathered (padded to the largest) and every rank rebuilds the full
    table.  Returns the handle of the full table; the local fragment table is destroyed."""
    import torch
    info = ctx.join_table_info(handle)
    counts = all_gather_counts(info.rows, world, device)
    stride = max(1, max(counts)) * info.row_bytes
    local = torch.empty(stride, dtype=torch.uint8, device=device)   # rows beyond the count are never read
    n = ctx.join_table_export(handle, local.data_ptr(), stride // info.row_bytes)
    assert n == info.rows
    gathered = all_gather_bytes(local, world)
    torch.cuda.synchronize(device)
    full = ctx.join_table_from_fragments(handle, gathered.data_ptr(), stride, counts)
    ctx.destroy_join_table(handle)
    return full


def or_merge_filter(rf, world: int, device) -> None:
    """Bloom OR-merge: all-gather the word arrays of every rank's filter (still in
    Building state) and OR them into the local one; bit exact by idempotence."""
    import torch
    words = torch.empty(rf.params.word_count * 8, dtype=torch.uint8, device=device)
    rf.read_words_into(words.data_ptr())          # device to device: the bits never visit the host
    gathered = all_gather_bytes(words, world)
    torch.cuda.synchronize(device)
    rf.or_device_words(gathered.data_ptr(), world)


def merge_partial_aggregate(plan, world: int, device, max_groups: int):
    """AggregateExec Partial -> Final across ranks: run the fused pipeline into a partial state,
    all-gather the states (padded to the largest) and merge them in rank order on every rank."""
    import torch
    cap = plan.partial_state_bytes(max_groups)
    state = torch.empty(cap, dtype=torch.uint8, device=device)      # the library writes count + entries
    nbytes, stats = plan.run_partial(state.data_ptr(), cap)
    sizes = all_gather_counts(nbytes, world, device)
    stride = (max(sizes) + 15) // 16 * 16
    gathered = all_gather_bytes(state[:stride].contiguous() if stride <= cap else torch.nn.functional.pad(state, (0, stride - cap)), world)
    torch.cuda.synchronize(device)
    return plan.merge_partials(gathered.data_ptr(), stride, world), stats
