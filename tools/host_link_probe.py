#!/usr/bin/env python
"""What the host side of the box can move: every rank copies pinned host memory to its GPU and back
at the same time (two streams per rank, all ranks released by a barrier), 1 GiB per direction and
repetition.  This is the ceiling of any host-buffer path at N ranks (bench.py's `e2e`): per-rank and
aggregate GB/s per direction, one JSON line from rank 0.

    python tools/host_link_probe.py                                   # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/host_link_probe.py
"""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = 1 << 30
    h_a = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    res = {}
    for mode in ("h2d", "d2h", "both"):
        best = None
        for _ in range(4):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_a.copy_(h_a, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_b.copy_(d_b, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        t = torch.tensor([n / best / 1e9], dtype=torch.float64, device=dev)
        if world > 1:
            allt = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allt, t)
        else:
            allt = [t]
        per = [float(x.item()) for x in allt]
        res[mode] = {"per_rank_gbs_per_direction": [round(x, 2) for x in per],
                     "aggregate_gbs_per_direction": round(sum(per), 2)}
    if rank == 0:
        print(json.dumps({"probe": "pinned H2D / D2H copies of 1 GiB, all ranks at once", "n_gpus": world, **res}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
