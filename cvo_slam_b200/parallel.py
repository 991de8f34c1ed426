"""Sharding of independent frame pairs across ranks (one process per GPU).

The alignment path has no exchange step: a pair touches only its two frames and its own
(R, T, ell), so pairs are partitioned across ranks with no collective on the data path
(SURVEY §8e).  The only communication is the final gather of the per-pair results
(~130 bytes per pair) and, in bench.py, a MAX all-reduce of the step time for reporting.
"""
from __future__ import annotations

import numpy as np


def partition(n_items, rank, world):
    """Static interleave: item p belongs to rank p % world (iteration counts vary between pairs
    by ~10x, an interleave spreads expensive neighbours; the per-GPU dynamic queue does the rest)."""
    return np.arange(rank, n_items, world, dtype=np.int64)


def partition_blocks(n_items, rank, world):
    """Contiguous blocks: rank r owns items [r * ceil(n / world), (r + 1) * ceil(n / world)).  In the
    verification pattern consecutive pairs share their fixed keyframe, so a block touches a compact set of
    frames and a rank selects points only for those (an interleave makes every rank select almost every
    frame); the same split as cvo_multi_align (csrc/multi.cu).  Iteration-count variance inside a block is
    absorbed by the per-GPU pair queue."""
    per = (n_items + world - 1) // world
    return np.arange(min(n_items, per * rank), min(n_items, per * (rank + 1)), dtype=np.int64)


def frames_needed(pairs, idx):
    """Sorted unique frame ids touched by the pairs `idx` of `pairs` (array [n, 2])."""
    pairs = np.asarray(pairs).reshape(-1, 2)
    return np.unique(pairs[idx].reshape(-1))


def gather_results(local, idx, n_total, group=None, dst=0):
    """Gathers per-pair result records (numpy structured array `local`, global indices `idx`) on
    rank `dst`; returns the full array there and None elsewhere.  Uses torch.distributed when it
    is initialised, otherwise returns the local array (single process)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        out = np.zeros(n_total, dtype=local.dtype)
        out[idx] = local
        return out
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    payload = (np.asarray(idx), local.tobytes(), str(local.dtype.descr))
    gathered = [None] * world if rank == dst else None
    dist.gather_object(payload, gathered, dst=dst, group=group)
    if rank != dst:
        return None
    out = np.zeros(n_total, dtype=local.dtype)
    for gi, raw, _ in gathered:
        out[gi] = np.frombuffer(raw, dtype=local.dtype)
    return out
