#!/usr/bin/env python
"""bench.py -- batched LZW encode + decode on B200 (BASELINE.json configs 3, 4 and 5).

    python bench.py --gpus N --steps K --warmup W [--impl reference]
                    [--workload config3|config4|config5|readme] [--scaling weak|strong]
                    [--endian le|be] [--streams S] [--single-process]

The default line is the contract's: config 3 (65,536 synthetic TIFF strips PER GPU, weak scaling).
`--workload config4` (4,096 GIF frames of 1 MiB, code sizes 2-8) and `--workload config5` (131,072
chunks of 64 KiB of lorem-like text, fixed 12-bit codes, `--endian le|be`) default to STRONG
scaling (the batch BASELINE.json names, sharded over the ranks by bytes); `--scaling` overrides.

One *step* = one pass of the hot path over the rank's batch: encode every stream, compact the
encoded streams into a dense buffer (prefix sum + gather), decode them back.  `value` =
uncompressed bytes of the whole job / step time (both directions are paid for every counted
byte), inputs resident in HBM.  `e2e` = the same step through the host-buffer C ABI
(slzw_encode_batch_host_dense + slzw_decode_batch_host) with pinned host memory, copies inside
the timed region.  `--single-process`: ONE process drives all N GPUs through slzw_multi_* (one
call per direction for the whole batch); only the end-to-end figure exists in that mode.

N > 1: one process per GPU (torchrun), no data-path collective (streams are independent);
time = max over ranks.  `--impl reference`: the reference's CPU algorithm (oracle/slzw_oracle.c,
the C restatement of salzweg -- the Rust crate cannot be built in this image) on all host
threads, on exactly the streams rank 0 of the GPU arm processes.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "uncompressed GB/s encode & decode, batched GIF/TIFF streams"
UNIT = "GB/s"
DEFAULT_STREAMS = {"config3": 65536, "config4": 4096, "config5": 131072}
DEFAULT_SCALING = {"config3": "weak", "config4": "strong", "config5": "strong"}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config3", choices=["config3", "config4", "config5", "readme"],
                    help="config3: the bench line of the contract; config4 / config5: the other multi-GPU configs of "
                         "BASELINE.json; readme: the rows of the reference's README (extra lines, not the contract's)")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"],
                    help="weak: --streams per GPU; strong: --streams in total, sharded by bytes "
                         "(default: weak for config3, strong for config4 / config5)")
    ap.add_argument("--streams", type=int, default=None,
                    help="streams (config3: 65,536 strips; config4: 4,096 frames; config5: 131,072 chunks)")
    ap.add_argument("--endian", default="le", choices=["le", "be"], help="config5: bit order of the 12-bit codes")
    ap.add_argument("--cpu-sample-mb", type=int, default=256, help="uncompressed MB in the CPU-baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--single-process", action="store_true",
                    help="one process, all --gpus devices through slzw_multi_* (end-to-end figure only)")
    args = ap.parse_args()
    if args.workload != "readme":
        args.streams = args.streams or DEFAULT_STREAMS[args.workload]
        args.scaling = args.scaling or DEFAULT_SCALING[args.workload]
    return args


# ---- workloads --------------------------------------------------------------------------------------
class Workload:
    """The streams one rank processes: buf/off (+ per-stream code sizes), GPU and oracle params."""


def make_workload(args, rank: int, world: int) -> Workload:
    from lzw_b200 import workloads as W
    from lzw_b200.types import Endianness, fixed_params, gif_params, tiff_params
    from oracle import oracle as O  # parameter structs of the checker / CPU arm only

    w = Workload()
    w.cs = None
    total_streams = args.streams * (world if args.scaling == "weak" else 1)
    if args.scaling == "weak":
        first, n = rank * args.streams, args.streams
    else:
        # strong: contiguous ranges of the one batch, balanced by bytes (slzw_partition_streams)
        import lzw_b200
        if args.workload == "config3":
            lens = np.empty(total_streams, dtype=np.uint64)
            W._wl().wl_tiff_strip_lens(W.SEED + 3, 0, total_streams, 8192, 57345, lens.ctypes.data)
        else:
            lens = np.full(total_streams, (1 << 20) if args.workload == "config4" else 65536, dtype=np.uint64)
        off_all = np.zeros(total_streams + 1, dtype=np.uint64)
        off_all[1:] = np.cumsum(lens)
        b = lzw_b200.partition_streams(off_all, world)
        first, n = int(b[rank]), int(b[rank + 1] - b[rank])
    if args.workload == "config3":
        w.buf, w.off = W.tiff_strips(n, first=first)
        w.params, w.oparams = tiff_params(), O.tiff()
        w.desc = ("config 3: synthetic TIFF strips, 8-64 KB, 4 entropy classes (random / photo walk / runs / Zipf "
                  "text), TIFF-style LZW (MSB-first, early change)")
    elif args.workload == "config4":
        w.buf, w.off, w.cs = W.gif_frames(n, first=first)
        w.params, w.oparams = gif_params(8), O.gif(8)
        w.desc = ("config 4: synthetic 1024x1024 8-bit-palette GIF frames, code size 2-8 per frame (runs + dither, "
                  "every 8th frame noise), GIF-style LZW (LSB-first)")
    else:
        big = args.endian == "be"
        w.buf, w.off = W.text_chunks(n, first=first)
        w.params = fixed_params(Endianness.BigEndian if big else Endianness.LittleEndian)
        w.oparams = O.fixed(big)
        w.desc = (f"config 5: synthetic lorem-like text in 64 KiB chunks, fixed 12-bit codes, "
                  f"{'MSB' if big else 'LSB'}-first")
    w.first, w.n, w.total_streams = first, n, total_streams
    return w


def workload_config(args, w: Workload, world: int, job_bytes: int):
    return {
        "workload": w.desc,
        "streams_total": int(w.total_streams),
        "streams_per_gpu": int(args.streams) if args.scaling == "weak" else None,
        "uncompressed_bytes_total": int(job_bytes),
        "scaling": args.scaling,
        "step": "encode + compact + decode of the whole batch",
        "generator": "SplitMix64, one sub-seed per stream (tools/wlgen/wlgen.c), seed 0x5A172E60 + config number",
        "l2": "inputs larger than L2 (no explicit flush)",
        "sharding": "streams sharded by rank, no collective",
    }


# ---- clocks ----------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def kernel_source_hash() -> str:
    """Identifies the kernels a committed ncu capture belongs to."""
    h = hashlib.sha256()
    for name in ("encode_kernels.cu", "decode_kernels.cu", "slzw_device.cuh"):
        with open(os.path.join(ROOT, "lzw_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def measured_traffic(kernel: str, workload: str, n_streams: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, from the committed
    `ncu --set full` capture of this same workload (profiles/r02_traffic.json).  None when there is
    no capture of this workload / stream count, or when the kernel sources changed since."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            t = json.load(f)
        if t.get("kernel_source_hash") != kernel_source_hash():
            return None
        e = t.get(workload, {}).get(kernel)
        if e and int(e.get("streams", -1)) == int(n_streams):
            return int(e["dram_bytes_read"]) + int(e["dram_bytes_write"])
    except (OSError, ValueError, KeyError):
        pass
    return None


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---- CPU baseline / reference arm ---------------------------------------------------------------------
def cpu_pass(w: Workload, lo: int, hi: int, threads: int):
    """One encode + decode pass of the oracle (restatement of salzweg) over streams [lo, hi) of the
    workload.  Returns (seconds_encode, seconds_decode, uncompressed bytes)."""
    from lzw_b200 import workloads as W
    from oracle import oracle as O
    off = (w.off[lo:hi + 1] - w.off[lo]).astype(np.uint64)
    buf = w.buf[int(w.off[lo]):int(w.off[hi])]
    cs = None if w.cs is None else w.cs[lo:hi]
    slots = W.encode_slots(off)
    t0 = time.perf_counter()
    out, out_len, status, _ = O.encode_batch(w.oparams, buf, off, slots, code_size=cs, threads=threads)
    t1 = time.perf_counter()
    # dense copy of the encoded streams, untimed (the CPU path writes each stream where it wants)
    dense_off = np.zeros(off.size, dtype=np.uint64)
    dense_off[1:] = np.cumsum(out_len)
    dense = np.empty(int(dense_off[-1]), dtype=np.uint8)
    for i in range(out_len.size):
        l = int(out_len[i])
        dense[int(dense_off[i]):int(dense_off[i]) + l] = out[int(slots[i]):int(slots[i]) + l]
    t2 = time.perf_counter()
    O.decode_batch(w.oparams, dense, dense_off, off, code_size=cs, threads=threads)
    t3 = time.perf_counter()
    return t1 - t0, t3 - t2, int(off[-1])


def streams_for_bytes(w: Workload, nbytes: int) -> int:
    rel = w.off - w.off[0]
    return int(max(1, min(w.n, np.searchsorted(rel, np.uint64(nbytes), side="right"))))


def cpu_baseline(args, w: Workload, threads: int):
    n = streams_for_bytes(w, args.cpu_sample_mb << 20)
    te, td, nbytes = cpu_pass(w, 0, n, threads)
    # single-threaded figure on a smaller sample (the north star asks for both)
    m = streams_for_bytes(w, 24 << 20)
    t1e, t1d, b1 = cpu_pass(w, 0, m, 1)
    return {
        "value": nbytes / (te + td) / 1e9,
        "unit": UNIT,
        "cores": threads,
        "single_thread": {"value": b1 / (t1e + t1d) / 1e9, "encode_gbs": b1 / t1e / 1e9,
                          "decode_gbs": b1 / t1d / 1e9, "sample_streams": m},
        "kind": "port",
        "sample": f"first {n} streams of the workload ({nbytes} uncompressed bytes), oracle/slzw_oracle.c "
                  f"(C restatement of salzweg; the Rust crate cannot be built here), one stream per task "
                  f"over {threads} host threads",
        "encode_gbs": nbytes / te / 1e9,
        "decode_gbs": nbytes / td / 1e9,
    }


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port) on all host cores, on exactly
    the streams rank 0 of the GPU arm processes (for N = 1: the whole job)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    w = make_workload(args, 0, world)
    nbytes = int(w.off[-1])
    job_bytes = nbytes * world if args.scaling == "weak" else None
    for _ in range(args.warmup):
        cpu_pass(w, 0, w.n, threads)
    te = td = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        a, b, _ = cpu_pass(w, 0, w.n, threads)
        te += a
        td += b
    wall = time.perf_counter() - t0
    value = nbytes * args.steps / (te + td) / 1e9
    sample = (f"the {w.n} streams of rank 0 ({nbytes} uncompressed bytes) per step"
              + ("" if world == 1 else f" -- one rank's share of the {world}-GPU job, the CPU arm is a rate")
              + f", oracle/slzw_oracle.c (C restatement of salzweg; no Rust toolchain to build the crate), "
                f"{threads} host threads")
    cfg = workload_config(args, w, world, job_bytes if job_bytes is not None else nbytes)
    if args.scaling == "strong" and world > 1:
        cfg["uncompressed_bytes_total"] = None  # only rank 0's shard is generated here
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": (te + td) / args.steps * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u8",
        "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "encode_gbs": nbytes * args.steps / te / 1e9,
                         "decode_gbs": nbytes * args.steps / td / 1e9},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall,
    }))


def bind_to_gpu_numa_node(index: int):
    """Runs this process on the CPUs of the NUMA node the GPU hangs off, so that the pinned host
    buffers of the end-to-end path (first touched after this call) are local to it.  Matters when
    several ranks share one host; returns the node or None."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(index)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except (OSError, ValueError, AttributeError, RuntimeError):
        pass
    return None


def host_link_probe(dev, nbytes: int = 1 << 30, sync=None):
    """Pinned H2D and D2H copies of `nbytes` each, running at the same time on two streams: what
    the host side of this process can move per direction while the other direction is busy (the
    ceiling of any host-buffer path; tools/e2e_probe.py measures several processes at once)."""
    import torch
    try:
        h_a = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        h_b = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    except RuntimeError:
        return None
    d_a = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_b = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    best = None
    for _ in range(3):
        torch.cuda.synchronize(dev)
        if sync:
            sync()  # all ranks copy at the same time: the ceiling the ranks' host paths share
        t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            d_a.copy_(h_a, non_blocking=True)
        with torch.cuda.stream(s2):
            h_b.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return nbytes / best / 1e9  # per direction, both directions busy


# ---- our arm ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import lzw_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; lzw_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    codec = lzw_b200.Codec(local_rank)
    from lzw_b200 import workloads as W
    w = make_workload(args, rank, world)
    params, buf, off, n = w.params, w.buf, w.off, w.n
    total = int(off[-1])
    slots = W.encode_slots(off)

    def i64(a):
        return torch.from_numpy(a.view(np.int64)).to(dev)

    t_in = torch.from_numpy(buf).to(dev)
    t_off = i64(off)
    t_slots = i64(slots)
    t_cs = torch.from_numpy(w.cs).to(dev) if w.cs is not None else None
    cs_ptr = t_cs.data_ptr() if t_cs is not None else 0
    t_enc = torch.empty(int(slots[-1]), dtype=torch.uint8, device=dev)
    t_enc_len = torch.zeros(n, dtype=torch.int64, device=dev)
    t_enc_st = torch.zeros(n, dtype=torch.int32, device=dev)
    t_enc_det = torch.zeros(n, dtype=torch.int32, device=dev)
    t_dense = torch.empty(int(slots[-1]), dtype=torch.uint8, device=dev)
    t_dense_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    t_dec = torch.empty(total, dtype=torch.uint8, device=dev)
    t_dec_len = torch.zeros(n, dtype=torch.int64, device=dev)
    t_dec_st = torch.zeros(n, dtype=torch.int32, device=dev)
    t_dec_det = torch.zeros(n, dtype=torch.int32, device=dev)

    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream

    def step(ev=None):
        if ev:
            ev[0].record(stream)
        codec.encode_batch_device(params, n, t_in.data_ptr(), t_off.data_ptr(), t_enc.data_ptr(),
                                  t_slots.data_ptr(), t_enc_len.data_ptr(), t_enc_st.data_ptr(),
                                  t_enc_det.data_ptr(), code_size_ptr=cs_ptr, stream=sp)
        if ev:
            ev[1].record(stream)
        codec.compact_device(t_enc.data_ptr(), t_slots.data_ptr(), t_enc_len.data_ptr(), n,
                             t_dense.data_ptr(), t_dense_off.data_ptr(), align=1, stream=sp)
        if ev:
            ev[2].record(stream)
        codec.decode_batch_device(params, n, t_dense.data_ptr(), t_dense_off.data_ptr(), t_dec.data_ptr(),
                                  t_off.data_ptr(), t_dec_len.data_ptr(), t_dec_st.data_ptr(),
                                  t_dec_det.data_ptr(), code_size_ptr=cs_ptr, stream=sp)
        if ev:
            ev[3].record(stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    launches0 = codec.kernel_launches
    start = torch.cuda.Event(enable_timing=True)
    stop = torch.cuda.Event(enable_timing=True)
    barrier()
    start.record(stream)
    for k in range(args.steps):
        step(evs[k])
    stop.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = codec.kernel_launches - launches0
    elapsed_ms = start.elapsed_time(stop)
    enc_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
    cmp_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
    dec_ms = float(np.mean([e[2].elapsed_time(e[3]) for e in evs]))

    # ---- parity properties at full size (outside the timed region) ----
    enc_st = t_enc_st.cpu().numpy()
    dec_st = t_dec_st.cpu().numpy()
    dec_len = t_dec_len.cpu().numpy()
    comp_total = int(t_dense_off[-1].item())
    lens = np.diff(off).astype(np.int64)
    round_trip_bytes_ok = bool(torch.equal(t_dec, t_in))  # SURVEY F1 streams still produce the right bytes
    parity = {
        "encode_status_ok": int((enc_st == 0).sum()),
        "decode_status_ok": int((dec_st == 0).sum()),
        "decode_len_matches": int((dec_len == lens).sum()),
        "round_trip_bytes_equal": round_trip_bytes_ok,
        "self_inconsistent_streams(F1)": int((dec_st != 0).sum()),
        "compressed_bytes": comp_total,
        "streams_deferred_to_exact_decoder": int(codec.last_deferred().size),
    }
    # oracle check of a sample (bytes, sizes, statuses) on every rank's first streams, gathered on rank 0
    from oracle import oracle as O
    m = streams_for_bytes(w, 24 << 20)
    o_out, o_len, o_st, _ = O.encode_batch(w.oparams, buf[: int(off[m])], off[: m + 1], slots[: m + 1],
                                           code_size=None if w.cs is None else w.cs[:m],
                                           threads=max(1, (os.cpu_count() or 1) // world))
    g_len = t_enc_len[:m].cpu().numpy().astype(np.uint64)
    g_out = t_enc[: int(slots[m])].cpu().numpy()
    same = bool(np.array_equal(g_len, o_len)) and bool(np.array_equal(enc_st[:m], o_st)) and all(
        np.array_equal(g_out[int(slots[i]):int(slots[i]) + int(o_len[i])],
                       o_out[int(slots[i]):int(slots[i]) + int(o_len[i])]) for i in range(m))
    parity["oracle_sample_streams"] = m
    parity["oracle_sample_byte_exact"] = same

    # ---- max over ranks; whole-job totals ----
    t = torch.tensor([elapsed_ms, enc_ms, cmp_ms, dec_ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(total), float(comp_total), float(n), float(same), float(round_trip_bytes_ok)],
                       dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        allt = [torch.zeros_like(tot) for _ in range(world)]
        dist.all_gather(allt, tot)
    else:
        allt = [tot]
    elapsed_ms, enc_ms, cmp_ms, dec_ms = [float(x) for x in t.tolist()]
    job_bytes = int(sum(float(x[0]) for x in allt))
    job_comp = int(sum(float(x[1]) for x in allt))
    parity["all_ranks_oracle_sample_byte_exact"] = all(bool(x[3]) for x in allt)
    parity["all_ranks_round_trip_bytes_equal"] = all(bool(x[4]) for x in allt)
    ms_per_step = elapsed_ms / args.steps
    value = job_bytes / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the host-buffer C ABI (pinned host memory) ----
    e2e = None
    if not args.no_e2e:
        # device memory of the kernel-only measurement is no longer needed
        del t_in, t_enc, t_dense, t_dec
        torch.cuda.empty_cache()
        link = host_link_probe(dev, sync=barrier)
        torch.cuda.empty_cache()
        dense_cap = comp_total + comp_total // 64 + (1 << 20)  # the compressed size is known by now
        pinned = True
        try:
            h_in = torch.from_numpy(buf).pin_memory().numpy()
            h_dense = torch.empty(dense_cap, dtype=torch.uint8).pin_memory().numpy()
            h_dec = torch.empty(total, dtype=torch.uint8).pin_memory().numpy()
        except RuntimeError:  # the host cannot page-lock this much: pageable buffers, copies are staged
            pinned = False
            h_in, h_dense, h_dec = buf, np.empty(dense_cap, dtype=np.uint8), np.empty(total, dtype=np.uint8)
        e2e_steps = max(1, min(args.steps, 3))

        def e2e_pass(hi, hd, ho):
            dense, doff, st, det = codec.encode_batch_dense(params, hi, off, code_size=w.cs, out=hd)
            dec, dlen, dst, ddet = codec.decode_batch(params, dense, doff, off, code_size=w.cs, out=ho)
            return dense, dec

        def timed(hi, hd, ho, steps):
            e2e_pass(hi, hd, ho)  # untimed: staging buffers of the host path are allocated on first use
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                dense, dec = e2e_pass(hi, hd, ho)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            te = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            return float(te.item()), dense, dec

        dt, dense, dec = timed(h_in, h_dense, h_dec, e2e_steps)
        h2d = total + dense.size + 3 * 8 * (n + 1)
        d2h = dense.size + total + 8 * (n + 1) + 8 * n + 4 * 4 * n
        e2e = {"value": job_bytes * e2e_steps / dt / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": e2e_steps, "round_trip_ok": bool(np.array_equal(dec[:total], buf)), "pinned": pinned,
               "numa_node_rank0": numa_node,
               "host_link_gbs_per_direction": link,
               "frac_of_host_link": (max(h2d, d2h) * e2e_steps / dt / 1e9 / link) if link else None,
               "note": "slzw_encode_batch_host_dense + slzw_decode_batch_host on pinned host buffers; every "
                       "byte crosses the bus by cudaMemcpyAsync inside the call (encode: one streaming launch "
                       "fed window by window); host_link = pinned H2D and D2H copies of 1 GiB running at the "
                       "same time, on all ranks at once (this rank's share)"}
        if pinned and world == 1 and total <= (4 << 30):
            # the same calls on pageable buffers (numpy arrays): every copy is staged by the driver
            p_dense = np.empty(dense_cap, dtype=np.uint8)
            p_dec = np.empty(total, dtype=np.uint8)
            dtp, _, decp = timed(buf, p_dense, p_dec, 1)
            e2e["pageable"] = {"value": job_bytes / dtp / 1e9, "unit": UNIT,
                               "round_trip_ok": bool(np.array_equal(decp[:total], buf))}

    if rank == 0:
        peak, peak_src = measured_peak()
        # the roofline describes one launch on one GPU: this rank's bytes
        kernels = {
            "encode": {"ms": enc_ms, "algorithmic_bytes": total + comp_total},
            "compact": {"ms": cmp_ms, "algorithmic_bytes": 2 * comp_total},
            # scheduler + fast kernel + exact kernel over the deferred streams
            "decode": {"ms": dec_ms, "algorithmic_bytes": total + comp_total},
        }
        for k in kernels.values():
            k["achieved_gbs"] = k["algorithmic_bytes"] / (k["ms"] * 1e-3) / 1e9
            k["frac_of_peak"] = k["achieved_gbs"] / peak
        dom = max(("encode", "decode"), key=lambda k: kernels[k]["ms"])
        wl_key = args.workload + ("_be" if args.workload == "config5" and args.endian == "be" else "")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, w, world, job_bytes),
            "encode_gbs": job_bytes / (enc_ms * 1e-3) / 1e9,
            "decode_gbs": job_bytes / (dec_ms * 1e-3) / 1e9,
            "compression_ratio": job_bytes / max(job_comp, 1),
            "roofline": {"bound": "hbm", "kernel": f"slzw_{dom}_kernel", "achieved": kernels[dom]["achieved_gbs"],
                         "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                         "frac": kernels[dom]["frac_of_peak"],
                         "traffic": measured_traffic(f"slzw_{dom}_kernel", wl_key, n),
                         "traffic_source": "profiles/r02_traffic.json (ncu --set full of this workload; null when the "
                                           "kernel sources changed since the capture)",
                         "algorithmic_bytes_per_launch": kernels[dom]["algorithmic_bytes"],
                         "launch_ms": kernels[dom]["ms"], "frac_of_nominal_8000": kernels[dom]["achieved_gbs"] / 8000.0,
                         "scope": "one launch on rank 0's GPU"},
            "kernels": kernels,
            "clocks": clocks, "gpu_launches": int(launches), "parity": parity,
        }
        if e2e:
            line["e2e"] = e2e
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(args, w, os.cpu_count() or 1)
        print(json.dumps(line))
    codec.close()
    if world > 1:
        dist.destroy_process_group()


def run_single_process(args):
    """One process, N GPUs, one call per direction for the whole batch (slzw_multi_*).  End to end
    only: host buffers in, host buffers out."""
    import torch

    import lzw_b200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; lzw_b200 has no CPU path")
    nd = min(args.gpus, torch.cuda.device_count())
    world_like = nd if args.scaling == "weak" else 1
    # the whole job is one host batch: weak scaling = N times the per-GPU stream count
    a2 = argparse.Namespace(**vars(args))
    a2.streams = args.streams * world_like
    a2.scaling = "strong"
    w = make_workload(a2, 0, 1)
    total = int(w.off[-1])
    mc = lzw_b200.MultiCodec(n_devices=nd)
    lens = np.diff(w.off)
    dense_cap = int((((lens + 3 + lens // 3838 + 1) * 12 + 7) // 8 + 1).sum())
    try:
        h_in = torch.from_numpy(w.buf).pin_memory().numpy()
        h_dense = torch.empty(dense_cap, dtype=torch.uint8).pin_memory().numpy()
        h_dec = torch.empty(total, dtype=torch.uint8).pin_memory().numpy()
        pinned = True
    except RuntimeError:
        h_in, h_dense, h_dec, pinned = w.buf, np.empty(dense_cap, dtype=np.uint8), np.empty(total, dtype=np.uint8), False

    def one():
        dense, doff, st, det = mc.encode_batch_dense(w.params, h_in, w.off, code_size=w.cs, out=h_dense)
        dec, dlen, dst, ddet = mc.decode_batch(w.params, dense, doff, w.off, code_size=w.cs, out=h_dec)
        return dense, doff, st, dec, dst

    for _ in range(max(1, min(args.warmup, 2))):
        one()
    sampler = ClockSampler(0)
    sampler.start()
    launches0 = mc.kernel_launches
    steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(steps):
        dense, doff, st, dec, dst = one()
    for d in range(nd):
        torch.cuda.synchronize(d)
    dt = time.perf_counter() - t0
    clocks = sampler.stop()
    # oracle sample
    from lzw_b200 import workloads as W
    from oracle import oracle as O
    m = streams_for_bytes(w, 24 << 20)
    sl = W.encode_slots(w.off[: m + 1])
    o_out, o_len, o_st, _ = O.encode_batch(w.oparams, w.buf[: int(w.off[m])], w.off[: m + 1], sl,
                                           code_size=None if w.cs is None else w.cs[:m], threads=os.cpu_count() or 1)
    same = bool(np.array_equal(np.diff(doff)[:m], o_len)) and all(
        np.array_equal(dense[int(doff[i]):int(doff[i + 1])], o_out[int(sl[i]):int(sl[i]) + int(o_len[i])])
        for i in range(m))
    value = total * steps / dt / 1e9
    n = w.n
    print(json.dumps({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": nd, "steps": steps, "warmup": args.warmup,
        "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "mode": "single process, slzw_multi_* (value IS the end-to-end figure)",
        "config": workload_config(a2, w, 1, total),
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": total + int(dense.size) + 3 * 8 * (n + 1),
                "d2h_bytes_per_step": int(dense.size) + total + 8 * (n + 1) + 8 * n + 16 * n, "pinned": pinned,
                "round_trip_ok": bool(np.array_equal(dec[:total], w.buf)),
                "note": "slzw_multi_encode_batch_host_dense + slzw_multi_decode_batch_host: one call per direction for "
                        "the whole batch, sharded by bytes over the devices, sizes gathered on the host"},
        "parity": {"oracle_sample_streams": m, "oracle_sample_byte_exact": same,
                   "encode_status_ok": int((st == 0).sum()), "decode_status_ok": int((dst == 0).sum())},
        "clocks": clocks, "gpu_launches": int(mc.kernel_launches - launches0),
    }))
    mc.close()


# ---- the reference's own README rows ----------------------------------------------------
# README.md:23-30 of the reference: single-thread criterion results on the indices of
# tokyo_128_colors.png (1024 x 684, code size 7) and on lorem_ipsum.txt, variable (GIF-style) and
# fixed 12-bit codes; inputs as lzw/benches/compare_crates.rs:4-16.  MiB/s of uncompressed bytes.
README_ROWS = {  # (input, codes, direction) -> published MiB/s (Ryzen 7 2700X, one thread)
    ("image", "variable", "encode"): 70, ("image", "fixed", "encode"): 120,
    ("image", "variable", "decode"): 200, ("image", "fixed", "decode"): 210,
    ("text", "variable", "encode"): 70, ("text", "fixed", "encode"): 85,
    ("text", "variable", "decode"): 200, ("text", "fixed", "decode"): 220,
}


def run_readme_rows(args):
    """One JSON line per README row: the GPU codec on one stream per call (what the reference's
    bench does, latency-bound on a GPU), on a device-resident batch of copies of the same stream
    (what the GPU codec is for), and the C port of the reference on one host thread."""
    import torch
    from PIL import Image

    import lzw_b200
    from lzw_b200.types import fixed_params, gif_params
    from oracle import oracle as O  # cpu_baseline leg of each row

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; lzw_b200 has no CPU path")
    dev = torch.device("cuda", 0)
    codec = lzw_b200.Codec(0)
    golden = os.path.join(ROOT, "tests", "golden")
    image = np.asarray(Image.open(os.path.join(golden, "tokyo_128_colors.png"))).reshape(-1).copy()
    with open(os.path.join(golden, "lorem_ipsum.txt"), "rb") as f:
        text = np.frombuffer(f.read(), dtype=np.uint8).copy()
    MiB = float(1 << 20)

    def i64(a):
        return torch.from_numpy(a.view(np.int64)).to(dev)

    for name, data, copies in (("image", image, 4096), ("text", text, 65536)):
        for codes, gp, op in (("variable", gif_params(7), O.gif(7)),
                              ("fixed", fixed_params(lzw_b200.Endianness.LittleEndian), O.fixed(False))):
            raw = data.tobytes()
            st, _, enc = codec.encode(gp, raw)
            ost, _, oenc = O.encode(op, raw)
            assert st == 0 and ost == 0 and enc == oenc, "GPU and oracle streams differ"
            # device-resident batch of `copies` copies
            n = copies
            off = np.arange(n + 1, dtype=np.uint64) * np.uint64(data.size)
            slot = (codec.encode_bound(gp, data.size) + 15) // 16 * 16
            slots = np.arange(n + 1, dtype=np.uint64) * np.uint64(slot)
            t_in = torch.from_numpy(data).to(dev).repeat(n)
            t_off, t_slots = i64(off), i64(slots)
            t_enc = torch.empty(int(slots[-1]), dtype=torch.uint8, device=dev)
            t_len = torch.zeros(n, dtype=torch.int64, device=dev)
            t_st = torch.zeros(n, dtype=torch.int32, device=dev)
            t_det = torch.zeros(n, dtype=torch.int32, device=dev)
            t_dec = torch.empty(n * data.size, dtype=torch.uint8, device=dev)
            t_dlen = torch.zeros(n, dtype=torch.int64, device=dev)
            t_dense = torch.empty(int(slots[-1]), dtype=torch.uint8, device=dev)
            t_doff = torch.zeros(n + 1, dtype=torch.int64, device=dev)

            def enc_batch():
                codec.encode_batch_device(gp, n, t_in.data_ptr(), t_off.data_ptr(), t_enc.data_ptr(),
                                          t_slots.data_ptr(), t_len.data_ptr(), t_st.data_ptr(), t_det.data_ptr())

            def dec_batch():
                codec.decode_batch_device(gp, n, t_dense.data_ptr(), t_doff.data_ptr(), t_dec.data_ptr(),
                                          t_off.data_ptr(), t_dlen.data_ptr(), t_st.data_ptr(), t_det.data_ptr())

            def time_gpu(fn, reps=5):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                return e0.elapsed_time(e1) / reps * 1e-3

            def time_host(fn, budget=1.5):
                fn()
                k, t0 = 0, time.perf_counter()
                while time.perf_counter() - t0 < budget:
                    fn()
                    k += 1
                return (time.perf_counter() - t0) / k

            t_be = time_gpu(enc_batch)
            enc_ok = bool((t_st == 0).all().item())
            codec.compact_device(t_enc.data_ptr(), t_slots.data_ptr(), t_len.data_ptr(), n, t_dense.data_ptr(),
                                 t_doff.data_ptr(), align=1)
            t_bd = time_gpu(dec_batch)
            dec_ok = bool((t_st == 0).all().item()) and bool(torch.equal(t_dec, t_in))
            for direction, t_batch, ok in (("encode", t_be, enc_ok), ("decode", t_bd, dec_ok)):
                if direction == "encode":
                    t_one = time_host(lambda: codec.encode(gp, raw))
                    t_cpu = time_host(lambda: O.encode(op, raw))
                else:
                    t_one = time_host(lambda: codec.decode(gp, enc, cap=len(raw)))
                    t_cpu = time_host(lambda: O.decode(op, enc, cap=len(raw)))
                published = README_ROWS[(name, codes, direction)]
                batch_mibs = n * data.size / t_batch / MiB
                print(json.dumps({
                    "metric": f"{direction}, {name} data, {codes} codes (reference README.md:27-30)",
                    "unit": "MiB/s", "higher_is_better": True, "dtype": "u8",
                    "data": "tests/golden (the reference's own bench inputs)",
                    "config": {"workload": f"{name}: {data.size} bytes, code size 7; batch = {n} copies, device-resident"},
                    "value": batch_mibs, "value_is": "GPU, batched",
                    "gpu_one_stream_per_call": len(raw) / t_one / MiB,
                    "published_reference_1_thread": published,
                    "vs_baseline": batch_mibs / published,
                    "cpu_baseline": {"value": len(raw) / t_cpu / MiB, "unit": "MiB/s", "cores": 1, "kind": "port",
                                     "sample": "the same single stream, oracle/slzw_oracle.c"},
                    "status_ok": ok, "byte_exact_vs_oracle": True,
                }), flush=True)
            del t_in, t_enc, t_dec, t_dense
            torch.cuda.empty_cache()
    codec.close()


def main():
    args = parse_args()
    if args.workload == "readme":
        return run_readme_rows(args)
    if args.impl == "reference":
        run_reference(args)
    elif args.single_process:
        run_single_process(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
