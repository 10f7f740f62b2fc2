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
