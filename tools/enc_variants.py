#!/usr/bin/env python
"""Times the encoder configurations (SLZW_ENC_CONFIG) on one batch and checks every one of them
against the first (status, sizes and bytes of all streams).

    python tools/enc_variants.py --streams 65536 --configs 0,10,11,12 [--workload config3|config4|config5]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=16384)
    ap.add_argument("--configs", default="0,10,11,12,13,14,15,16,17,18")
    ap.add_argument("--workload", default="config3")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch

    import lzw_b200
    from lzw_b200 import workloads as W
    from lzw_b200.types import fixed_params, gif_params, tiff_params

    cs = None
    if args.workload == "config4":
        buf, off, cs = W.gif_frames(args.streams)
        params = gif_params(8)
    elif args.workload == "config5":
        buf, off = W.text_chunks(args.streams)
        params = fixed_params(lzw_b200.Endianness.LittleEndian)
    else:
        buf, off = W.tiff_strips(args.streams)
        params = tiff_params()
    n = off.size - 1
    slots = W.encode_slots(off)
    dev = torch.device("cuda:0")
    d_in = torch.from_numpy(buf).to(dev)
    d_off = torch.from_numpy(off.view(np.int64)).to(dev)
    d_slots = torch.from_numpy(slots.view(np.int64)).to(dev)
    d_cs = torch.from_numpy(cs).to(dev) if cs is not None else None
    d_out = torch.empty(int(slots[-1]), dtype=torch.uint8, device=dev)
    d_len = torch.zeros(n, dtype=torch.int64, device=dev)
    d_st = torch.zeros(n, dtype=torch.int32, device=dev)
    d_det = torch.zeros(n, dtype=torch.int32, device=dev)
    d_dense = torch.empty(int(slots[-1]), dtype=torch.uint8, device=dev)
    d_doff = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    ref = None
    for cfg in [int(c) for c in args.configs.split(",")]:
        os.environ["SLZW_ENC_CONFIG"] = str(cfg)
        codec = lzw_b200.Codec(0)
        times = []
        for rep in range(args.reps + 1):
            d_out.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            codec.encode_batch_device(params, n, d_in.data_ptr(), d_off.data_ptr(), d_out.data_ptr(),
                                      d_slots.data_ptr(), d_len.data_ptr(), d_st.data_ptr(), d_det.data_ptr(),
                                      d_cs.data_ptr() if d_cs is not None else 0,
                                      stream=torch.cuda.current_stream().cuda_stream)
            e1.record()
            torch.cuda.synchronize()
            if rep:
                times.append(e0.elapsed_time(e1))
        codec.compact_device(d_out.data_ptr(), d_slots.data_ptr(), d_len.data_ptr(), n, d_dense.data_ptr(),
                             d_doff.data_ptr(), 1, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        total = int(d_doff[-1].item())
        res = (d_len.clone(), d_st.clone(), d_det.clone(), d_dense[:total].clone())
        ok = True
        if ref is None:
            ref = res
        else:
            ok = all(torch.equal(a, b) for a, b in zip(ref, res))
        ms = float(np.median(times))
        import hashlib
        digest = hashlib.sha256(d_dense[:total].cpu().numpy()).hexdigest()[:16]  # compares builds (SLZW_LIB)
        shares = codec.last_encode_shares()
        print(json.dumps({"config": cfg, "workload": args.workload, "streams": n, "ms": round(ms, 3),
                          "GBps_uncompressed": round(buf.size / ms / 1e6, 2), "compressed": total,
                          "errors": int((d_st != 0).sum().item()), "matches_first": ok, "sha256_16": digest,
                          "GBps_by_kind[tmem_warp,smem_warp,smem_lanes,global_lanes]": [round(float(x) / ms / 1e6, 2) for x in shares]}), flush=True)
        codec.close()


if __name__ == "__main__":
    main()
