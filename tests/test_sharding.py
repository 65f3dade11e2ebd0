"""Host-side multi-GPU logic: partition of a batch by stream and the size gather, covered with a
world_size-2 gloo run on CPU (no codec calls: the per-rank 'encoded sizes' are computed by the
oracle, which is the checker here)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from lzw_b200 import sharding
from oracle import oracle as O
from tests import cases as T


def test_partition_covers_every_stream_once():
    rng = np.random.default_rng(3)
    for n in (0, 1, 2, 7, 100, 1000):
        lens = rng.integers(0, 5000, size=n)
        off = np.zeros(n + 1, dtype=np.uint64)
        off[1:] = np.cumsum(lens)
        for world in (1, 2, 3, 4, 8):
            b = sharding.partition(off, world)
            assert b[0] == 0 and b[-1] == n and np.all(np.diff(b) >= 0)
            if n >= 8 * world and off[-1] > 0:
                sizes = [int(off[b[r + 1]] - off[b[r]]) for r in range(world)]
                assert max(sizes) - min(sizes) <= 2 * int(lens.max()), "ranges are byte-balanced"


def test_c_abi_partition_matches_the_host_model():
    """slzw_partition_streams (what slzw_multi_* shards with; pure host code, runs without a GPU)
    against sharding.partition."""
    import lzw_b200
    rng = np.random.default_rng(5)
    for n in (0, 1, 2, 7, 100, 1000, 5000):
        lens = rng.integers(0, 70000, size=n)
        off = np.zeros(n + 1, dtype=np.uint64)
        off[1:] = np.cumsum(lens)
        off += np.uint64(12345)  # absolute offsets need not start at 0
        for world in (1, 2, 3, 4, 8):
            assert np.array_equal(lzw_b200.partition_streams(off, world).astype(np.int64),
                                  sharding.partition(off, world)), (n, world)


def test_shard_views_are_rebased():
    buf, off = T.make_batch(11, 40, 255)
    pieces = []
    for r in range(4):
        sub, sub_off, first = sharding.shard(buf, off, 4, r)
        assert sub_off[0] == 0 and sub.size == int(sub_off[-1])
        pieces.append(sub)
    assert np.array_equal(np.concatenate(pieces), buf)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    buf, off = T.make_batch(5, 64, 255, max_len=3000)
    sub, sub_off, first = sharding.shard(buf, off, world, rank)
    slots = np.zeros(sub_off.size, dtype=np.uint64)
    slots[1:] = np.cumsum([O.encode_bound(int(l)) for l in np.diff(sub_off)])
    _, out_len, status, _ = O.encode_batch(O.tiff(), sub, sub_off, slots)
    all_len, all_status, place = sharding.gather_sizes(out_len, status)
    q.put((rank, first, all_len.tolist(), all_status.tolist(), place.tolist()))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gather_matches_single_rank():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    buf, off = T.make_batch(5, 64, 255, max_len=3000)
    slots = np.zeros(off.size, dtype=np.uint64)
    slots[1:] = np.cumsum([O.encode_bound(int(l)) for l in np.diff(off)])
    _, want_len, want_status, _ = O.encode_batch(O.tiff(), buf, off, slots)
    for rank, first, all_len, all_status, place in results:
        assert all_len == want_len.tolist() and all_status == want_status.tolist()
        assert place[-1] == int(want_len.sum())
