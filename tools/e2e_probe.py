"""Times the host-buffer entry points (pinned memory) for several pipeline chunk sizes, and the raw
PCIe copy rates (debug aid for the e2e figure of bench.py)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import lzw_b200
    from lzw_b200 import workloads as W
    from lzw_b200.types import tiff_params

    streams = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    buf, off = W.tiff_strips(streams)
    total = int(off[-1])
    slots = W.encode_slots(off)
    h_in = torch.from_numpy(buf).pin_memory()
    d = torch.empty(total, dtype=torch.uint8, device="cuda")
    for name, fn in (("H2D", lambda: d.copy_(h_in, non_blocking=True)), ("D2H", lambda: h_in.copy_(d, non_blocking=True))):
        fn(); torch.cuda.synchronize()
        t = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t
        print(f"{name} {total / dt / 1e9:.1f} GB/s pinned, {total} bytes", flush=True)
    h_in = h_in.numpy()
    h_dense = torch.empty(int(slots[-1]), dtype=torch.uint8).pin_memory().numpy()
    h_dec = torch.empty(total, dtype=torch.uint8).pin_memory().numpy()
    p = tiff_params()
    for chunk, zc in ((0, 1), (0, 2), (0, 0)):
        os.environ.pop("SLZW_HOST_CHUNK_BYTES", None)
        if chunk:
            os.environ["SLZW_HOST_CHUNK_BYTES"] = str(chunk)
        os.environ["SLZW_HOST_ZERO_COPY"] = str(zc)
        codec = lzw_b200.Codec(0)
        for it in range(2):
            t0 = time.perf_counter()
            dense, doff, st, det = codec.encode_batch_dense(p, h_in, off, out=h_dense)
            t1 = time.perf_counter()
            dec, dlen, dst, ddet = codec.decode_batch(p, dense, doff, off, out=h_dec)
            t2 = time.perf_counter()
        print(f"chunk {chunk >> 20} MiB (0 = defaults), zero-copy input {zc}: encode {1e3 * (t1 - t0):.1f} ms, decode {1e3 * (t2 - t1):.1f} ms, "
              f"e2e {total / (t2 - t0) / 1e9:.2f} GB/s, ok={bool(np.array_equal(dec[:total], buf))}", flush=True)
        codec.close()


if __name__ == "__main__":
    main()
