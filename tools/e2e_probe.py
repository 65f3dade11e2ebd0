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
    # settings: "name=ENV1=val,ENV2=val" arguments after the stream count; default = the shipped path
    settings = [a.split("=", 1) for a in sys.argv[2:]] or [["default", ""]]
    knobs = ("SLZW_HOST_CHUNK_BYTES", "SLZW_HOST_ZERO_COPY", "SLZW_HOST_ENC_HILL", "SLZW_HOST_ENC_STREAM",
             "SLZW_HOST_WINDOW_MB", "SLZW_HOST_STREAM_SMS")
    for name, envs in settings:
        for k in knobs:
            os.environ.pop(k, None)
        for kv in filter(None, envs.split(",")):
            k, v = kv.split("=", 1)
            os.environ[k] = v
        codec = lzw_b200.Codec(0)
        best_e = best_d = 1e9
        for it in range(4):
            t0 = time.perf_counter()
            dense, doff, st, det = codec.encode_batch_dense(p, h_in, off, out=h_dense)
            t1 = time.perf_counter()
            dec, dlen, dst, ddet = codec.decode_batch(p, dense, doff, off, out=h_dec)
            t2 = time.perf_counter()
            if it:
                best_e, best_d = min(best_e, t1 - t0), min(best_d, t2 - t1)
        print(f"{name} [{envs}]: encode {1e3 * best_e:.1f} ms, decode {1e3 * best_d:.1f} ms, "
              f"e2e {total / (best_e + best_d) / 1e9:.2f} GB/s, ok={bool(np.array_equal(dec[:total], buf))}", flush=True)
        codec.close()


if __name__ == "__main__":
    main()
