"""Sharding a batch of independent streams across the GPUs of one box.

Streams carry no cross-stream state (every codec loop of the reference keeps its state in
locals: encoder.rs:273-346, decoder.rs:174-290), so a batch shards by stream with no data-path
collective.  The only cross-rank step is bookkeeping: gathering the per-stream sizes / statuses
on the host so that the shards can be placed in one output buffer.
"""
from __future__ import annotations

import numpy as np


def partition(off: np.ndarray, world: int) -> np.ndarray:
    """Contiguous stream ranges balanced by uncompressed bytes.

    off: n + 1 offsets.  Returns world + 1 stream indices b with rank r owning streams
    [b[r], b[r+1]); every stream belongs to exactly one rank, ranks may be empty when n < world."""
    off = np.asarray(off, dtype=np.uint64)
    n = off.size - 1
    if world < 1:
        raise ValueError("world must be >= 1")
    total = int(off[-1] - off[0])
    bounds = np.zeros(world + 1, dtype=np.int64)
    bounds[world] = n
    rel = (off - off[0]).astype(np.uint64)
    for r in range(1, world):
        target = total * r // world
        # first stream whose start is at or past the target byte
        b = int(np.searchsorted(rel, np.uint64(target), side="left"))
        bounds[r] = min(max(b, int(bounds[r - 1])), n)
    return bounds


def shard(buf: np.ndarray, off: np.ndarray, world: int, rank: int):
    """This rank's streams as (buffer view, rebased offsets, first stream index)."""
    b = partition(off, world)
    lo, hi = int(b[rank]), int(b[rank + 1])
    o = np.asarray(off, dtype=np.uint64)
    return buf[int(o[lo]):int(o[hi])], (o[lo:hi + 1] - o[lo]).astype(np.uint64), lo


def gather_sizes(local_len: np.ndarray, local_status: np.ndarray, group=None):
    """Host-side gather of per-stream sizes / statuses over torch.distributed (any backend).
    Returns (all_len, all_status, placement offsets of every stream in the merged output)."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    parts = [None] * world
    dist.all_gather_object(parts, (np.asarray(local_len, dtype=np.uint64),
                                   np.asarray(local_status, dtype=np.uint32)), group=group)
    all_len = np.concatenate([p[0] for p in parts]) if parts else np.zeros(0, dtype=np.uint64)
    all_status = np.concatenate([p[1] for p in parts]) if parts else np.zeros(0, dtype=np.uint32)
    place = np.zeros(all_len.size + 1, dtype=np.uint64)
    place[1:] = np.cumsum(all_len)
    return all_len, all_status, place
