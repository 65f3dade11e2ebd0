"""Times the TIFF predictor kernels (difference / accumulate) on device-resident strips and prints
their HBM traffic rate (one read + one write of every byte).  Usage: python tools/predictor_bw.py"""
import json
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lzw_b200  # noqa: E402


def main():
    codec = lzw_b200.Codec(0)
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    out = []
    for name, n, lo, hi, row_bytes, spp in (("65536 strips 8-64 KB, RGB rows of 3072 B", 65536, 8192, 65536, 3072, 3),
                                            ("65536 strips 8-64 KB, grey rows of 1024 B", 65536, 8192, 65536, 1024, 1),
                                            ("4096 strips of 1 MiB, RGBA rows of 4096 B", 4096, 1 << 20, (1 << 20) + 1, 4096, 4)):
        lens = rng.integers(lo, hi, size=n).astype(np.uint64)
        off = np.zeros(n + 1, dtype=np.uint64)
        off[1:] = np.cumsum(lens)
        total = int(off[-1])
        t = torch.randint(0, 256, (total + 64,), dtype=torch.uint8, device=dev)
        t_off = torch.from_numpy(off.view(np.int64)).to(dev)
        for direction, label in ((codec.DIFFERENCE, "difference"), (codec.ACCUMULATE, "accumulate")):
            for _ in range(3):
                codec.tiff_predictor_device(direction, t.data_ptr(), t_off.data_ptr(), n, row_bytes, spp)
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            reps = 10
            ev[0].record()
            for _ in range(reps):
                codec.tiff_predictor_device(direction, t.data_ptr(), t_off.data_ptr(), n, row_bytes, spp)
            ev[1].record()
            torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1]) / reps
            out.append({"workload": name, "kernel": label, "bytes": total, "ms": round(ms, 4),
                        "GBps_traffic": round(2 * total / ms / 1e6, 1)})
            print(json.dumps(out[-1]), flush=True)
    codec.close()


if __name__ == "__main__":
    main()
