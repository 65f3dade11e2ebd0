"""Runs a few encode / compact / decode passes on a small config-3 batch (for ncu captures)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=2048)
    ap.add_argument("--passes", type=int, default=2)
    ap.add_argument("--what", default="all", choices=["all", "encode", "decode"])
    ap.add_argument("--cls", type=int, default=-1, help="keep only streams of entropy class cls (i mod 4)")
    ap.add_argument("--config", type=int, default=3, choices=[3, 4, 5],
                    help="BASELINE config: 3 TIFF strips, 4 GIF frames (1 MiB, cs 2..8), 5 fixed 12-bit text chunks")
    args = ap.parse_args()
    import torch

    import lzw_b200
    from lzw_b200 import workloads as W
    from lzw_b200.types import Endianness, fixed_params, gif_params, tiff_params

    dev = torch.device("cuda:0")
    codec = lzw_b200.Codec(0)
    p = tiff_params()
    cs = None
    if args.config == 4:
        p = gif_params(8)
        buf, off, cs = W.gif_frames(args.streams)
    elif args.config == 5:
        p = fixed_params(Endianness.LittleEndian)
        buf, off = W.text_chunks(args.streams, corpus_bytes=16 << 20)
    else:
        buf, off = W.tiff_strips(args.streams)
    if args.cls >= 0:
        idx = np.arange(args.cls, args.streams, 4)
        parts = [buf[int(off[i]):int(off[i + 1])] for i in idx]
        off = np.zeros(len(parts) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([p.size for p in parts])
        buf = np.concatenate(parts)
    slots = W.encode_slots(off)
    n = off.size - 1
    i64 = lambda a: torch.from_numpy(a.view(np.int64)).to(dev)
    t_in, t_off, t_slots = torch.from_numpy(buf).to(dev), i64(off), i64(slots)
    t_enc = torch.empty(int(slots[-1]), dtype=torch.uint8, device=dev)
    t_len = torch.zeros(n, dtype=torch.int64, device=dev)
    t_st = torch.zeros(n, dtype=torch.int32, device=dev)
    t_det = torch.zeros(n, dtype=torch.int32, device=dev)
    t_dense = torch.empty(int(slots[-1]), dtype=torch.uint8, device=dev)
    t_doff = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    t_dec = torch.empty(buf.size, dtype=torch.uint8, device=dev)
    t_cs = torch.from_numpy(cs).to(dev) if cs is not None else None
    cs_ptr = t_cs.data_ptr() if t_cs is not None else 0
    for _ in range(args.passes):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        codec.encode_batch_device(p, n, t_in.data_ptr(), t_off.data_ptr(), t_enc.data_ptr(), t_slots.data_ptr(),
                                  t_len.data_ptr(), t_st.data_ptr(), t_det.data_ptr(), code_size_ptr=cs_ptr)
        ev[1].record()
        codec.compact_device(t_enc.data_ptr(), t_slots.data_ptr(), t_len.data_ptr(), n, t_dense.data_ptr(),
                             t_doff.data_ptr())
        ev[2].record()
        if args.what != "encode":
            codec.decode_batch_device(p, n, t_dense.data_ptr(), t_doff.data_ptr(), t_dec.data_ptr(), t_off.data_ptr(),
                                      t_len.data_ptr(), t_st.data_ptr(), t_det.data_ptr(), code_size_ptr=cs_ptr)
        ev[3].record()
        torch.cuda.synchronize()
        print(f"bytes={buf.size} encode_ms={ev[0].elapsed_time(ev[1]):.3f} compact_ms={ev[1].elapsed_time(ev[2]):.3f} "
              f"decode_ms={ev[2].elapsed_time(ev[3]):.3f}")
    if args.what != "encode":
        print("round trip equal:", bool(torch.equal(t_dec, t_in)))
        d = codec.last_deferred()
        st = t_st.cpu().numpy()
        ln = t_len.cpu().numpy()
        print("deferred:", d.size, [(int(i), int(i) % 4, int(off[i + 1] - off[i]), int(ln[i]), int(st[i])) for i in d[:16]])
    codec.close()


if __name__ == "__main__":
    main()
