#!/usr/bin/env python
"""bench.py -- the headline benchmark: batched TIFF-style LZW encode + decode (BASELINE config 3).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--streams S]

One *step* = one pass of the hot path over one batch of synthetic TIFF strips: encode every
strip, compact the compressed strips into a dense buffer (prefix sum + gather), decode them
back.  `value` = uncompressed bytes of the batch / step time (so both directions are paid for
every counted byte), inputs resident in HBM.  `e2e` = the same step through the host-buffer C
ABI (slzw_{encode,decode}_batch_host) with pinned host memory, copies inside the timed region.

N > 1: one process per GPU (torchrun), every rank owns its own batch of the same shape (weak
scaling, no data-path collective: streams are independent); time = max over ranks.
"""
from __future__ import annotations

import argparse
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=65536, help="TIFF strips per GPU (config 3: 65,536)")
    ap.add_argument("--cpu-sample", type=int, default=4096, help="strips in the CPU-baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="config3", choices=["config3", "readme"],
                    help="config3: the bench line of the contract; readme: the rows of the reference's README "
                         "(tokyo image / lorem text, variable and fixed codes) -- extra lines, not the contract's")
    return ap.parse_args()


def workload_config(args, n_streams, total_bytes):
    return {
        "workload": f"config 3: {n_streams} synthetic TIFF strips per GPU, 8-64 KB, 4 entropy classes "
                    "(random / photo walk / runs / Zipf text), TIFF-style LZW (MSB-first, early change)",
        "streams_per_gpu": n_streams,
        "uncompressed_bytes_per_gpu": int(total_bytes),
        "step": "encode + compact + decode of the whole batch",
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


def measured_traffic(kernel: str, n_streams: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, from the committed
    `ncu --set full` capture of this same workload (profiles/r01_traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            t = json.load(f)
        e = t.get(kernel)
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
def cpu_pass(buf, off, threads):
    """One encode + decode pass of the oracle (restatement of salzweg) over a batch.
    Returns (seconds_encode, seconds_decode)."""
    from lzw_b200 import workloads as W
    from oracle import oracle as O
    slots = W.encode_slots(off)
    t0 = time.perf_counter()
    out, out_len, status, _ = O.encode_batch(O.tiff(), buf, off, slots, threads=threads)
    t1 = time.perf_counter()
    # dense copy of the encoded strips, untimed (the CPU path writes each strip where it wants)
    dense_off = np.zeros(off.size, dtype=np.uint64)
    dense_off[1:] = np.cumsum(out_len)
    dense = np.empty(int(dense_off[-1]), dtype=np.uint8)
    for i in range(out_len.size):
        l = int(out_len[i])
        dense[int(dense_off[i]):int(dense_off[i]) + l] = out[int(slots[i]):int(slots[i]) + l]
    t2 = time.perf_counter()
    O.decode_batch(O.tiff(), dense, dense_off, off, threads=threads)
    t3 = time.perf_counter()
    return t1 - t0, t3 - t2


def cpu_baseline(buf, off, sample_streams, threads):
    n = min(sample_streams, off.size - 1)
    sub_off = off[: n + 1] - off[0]
    sub = buf[int(off[0]):int(off[n])]
    te, td = cpu_pass(sub, sub_off, threads)
    nbytes = int(sub_off[-1])
    # single-threaded figure on a smaller sample (the north star asks for both)
    m = min(256, n)
    t1e, t1d = cpu_pass(buf[int(off[0]):int(off[m])], off[: m + 1] - off[0], 1)
    b1 = int(off[m] - off[0])
    return {
        "value": nbytes / (te + td) / 1e9,
        "unit": UNIT,
        "cores": threads,
        "single_thread": {"value": b1 / (t1e + t1d) / 1e9, "encode_gbs": b1 / t1e / 1e9,
                          "decode_gbs": b1 / t1d / 1e9, "sample_streams": m},
        "kind": "port",
        "sample": f"first {n} strips of the workload ({nbytes} uncompressed bytes), oracle/slzw_oracle.c "
                  f"(C restatement of salzweg; the Rust crate cannot be built here), one stream per task "
                  f"over {threads} host threads",
        "encode_gbs": nbytes / te / 1e9,
        "decode_gbs": nbytes / td / 1e9,
    }


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from lzw_b200 import workloads as W
    threads = os.cpu_count() or 1
    n = min(args.cpu_sample, args.streams)
    buf, off = W.tiff_strips(n)
    nbytes = int(off[-1])
    for _ in range(args.warmup):
        cpu_pass(buf, off, threads)
    te = td = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        a, b = cpu_pass(buf, off, threads)
        te += a
        td += b
    wall = time.perf_counter() - t0
    value = nbytes * args.steps / (te + td) / 1e9
    sample = (f"{n} strips ({nbytes} uncompressed bytes) of config 3 per step, oracle/slzw_oracle.c "
              f"(C restatement of salzweg; no Rust toolchain to build the crate), {threads} host threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": (te + td) / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic", "config": workload_config(args, n, nbytes),
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


# ---- our arm ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import lzw_b200
    from lzw_b200 import workloads as W
    from lzw_b200.types import tiff_params

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
    params = tiff_params()

    # ---- synthetic batch of this rank (weak scaling: same shape, different seed per rank) ----
    buf, off = W.tiff_strips(args.streams, seed=W.SEED + 3 + 1000 * rank)
    n = off.size - 1
    total = int(off[-1])
    slots = W.encode_slots(off)

    def i64(a):
        return torch.from_numpy(a.view(np.int64)).to(dev)

    t_in = torch.from_numpy(buf).to(dev)
    t_off = i64(off)
    t_slots = i64(slots)
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
                                  t_enc_det.data_ptr(), stream=sp)
        if ev:
            ev[1].record(stream)
        codec.compact_device(t_enc.data_ptr(), t_slots.data_ptr(), t_enc_len.data_ptr(), n,
                             t_dense.data_ptr(), t_dense_off.data_ptr(), align=1, stream=sp)
        if ev:
            ev[2].record(stream)
        codec.decode_batch_device(params, n, t_dense.data_ptr(), t_dense_off.data_ptr(), t_dec.data_ptr(),
                                  t_off.data_ptr(), t_dec_len.data_ptr(), t_dec_st.data_ptr(),
                                  t_dec_det.data_ptr(), stream=sp)
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
    # oracle check of a sample (bytes, sizes, statuses), rank 0 only
    if rank == 0:
        from oracle import oracle as O
        m = min(512, n)
        o_out, o_len, o_st, _ = O.encode_batch(O.tiff(), buf[: int(off[m])], off[: m + 1], slots[: m + 1],
                                               threads=os.cpu_count() or 1)
        g_len = t_enc_len[:m].cpu().numpy().astype(np.uint64)
        g_out = t_enc[: int(slots[m])].cpu().numpy()
        same = bool(np.array_equal(g_len, o_len)) and all(
            np.array_equal(g_out[int(slots[i]):int(slots[i]) + int(o_len[i])],
                           o_out[int(slots[i]):int(slots[i]) + int(o_len[i])]) for i in range(m))
        parity["oracle_sample_streams"] = m
        parity["oracle_sample_byte_exact"] = same

    # ---- max over ranks ----
    t = torch.tensor([elapsed_ms, enc_ms, cmp_ms, dec_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, enc_ms, cmp_ms, dec_ms = [float(x) for x in t.tolist()]
    ms_per_step = elapsed_ms / args.steps
    value = world * total / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the host-buffer C ABI (pinned host memory) ----
    e2e = None
    if not args.no_e2e:
        # device memory of the kernel-only measurement is no longer needed
        del t_in, t_enc, t_dense, t_dec
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
        h2d = d2h = 0
        # one untimed pass: staging buffers of the host path are allocated on first use
        dense, doff, st, det = codec.encode_batch_dense(params, h_in, off, out=h_dense)
        codec.decode_batch(params, dense, doff, off, out=h_dec)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            dense, doff, st, det = codec.encode_batch_dense(params, h_in, off, out=h_dense)
            dec, dlen, dst, ddet = codec.decode_batch(params, dense, doff, off, out=h_dec)
            h2d += total + dense.size + 3 * 8 * (n + 1)
            d2h += dense.size + total + 8 * (n + 1) + 8 * n + 4 * 4 * n
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        te = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dt = float(te.item())
        e2e = {"value": world * total * e2e_steps / dt / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": h2d // e2e_steps, "d2h_bytes_per_step": d2h // e2e_steps,
               "steps": e2e_steps, "round_trip_ok": bool(np.array_equal(dec[:total], buf)), "pinned": pinned,
               "numa_node_rank0": numa_node,
               "note": "slzw_encode_batch_host_dense + slzw_decode_batch_host on pinned host buffers; the "
                       "encoder reads its pinned input in place over PCIe (counted in h2d_bytes_per_step), "
                       "everything else is cudaMemcpyAsync inside the call"}

    if rank == 0:
        peak, peak_src = measured_peak()
        comp = parity["compressed_bytes"]
        kernels = {
            "encode": {"ms": enc_ms, "algorithmic_bytes": total + comp},
            "compact": {"ms": cmp_ms, "algorithmic_bytes": 2 * comp},
            # scheduler + fast kernel + exact kernel over the deferred streams
            "decode": {"ms": dec_ms, "algorithmic_bytes": total + comp},
        }
        for k in kernels.values():
            k["achieved_gbs"] = k["algorithmic_bytes"] / (k["ms"] * 1e-3) / 1e9
            k["frac_of_peak"] = k["achieved_gbs"] / peak
        dom = max(("encode", "decode"), key=lambda k: kernels[k]["ms"])
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, n, total),
            "encode_gbs": world * total / (enc_ms * 1e-3) / 1e9,
            "decode_gbs": world * total / (dec_ms * 1e-3) / 1e9,
            "roofline": {"bound": "hbm", "kernel": f"slzw_{dom}_kernel", "achieved": kernels[dom]["achieved_gbs"],
                         "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                         "frac": kernels[dom]["frac_of_peak"],
                         "traffic": measured_traffic(f"slzw_{dom}_kernel", n),
                         "algorithmic_bytes_per_launch": kernels[dom]["algorithmic_bytes"],
                         "launch_ms": kernels[dom]["ms"], "frac_of_nominal_8000": kernels[dom]["achieved_gbs"] / 8000.0},
            "kernels": kernels,
            "clocks": clocks, "gpu_launches": int(launches), "parity": parity,
        }
        if e2e:
            line["e2e"] = e2e
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(buf, off, args.cpu_sample, os.cpu_count() or 1)
        print(json.dumps(line))
    codec.close()
    if world > 1:
        dist.destroy_process_group()


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
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
