"""Encode kernel alone (CUDA events) with its input in device memory and in pinned host memory read
in place over PCIe: is the in-place read what the host path's encode call waits for?"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import lzw_b200
    from lzw_b200 import workloads as W
    from lzw_b200.types import tiff_params

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    buf, off = W.tiff_strips(n)
    slots = W.encode_slots(off)
    dev = torch.device("cuda:0")
    h_in = torch.from_numpy(buf).pin_memory()
    d_in = h_in.to(dev)
    t_off = torch.from_numpy(off.view(np.int64)).to(dev)
    t_slots = torch.from_numpy(slots.view(np.int64)).to(dev)
    t_out = torch.empty(int(slots[-1]), dtype=torch.uint8, device=dev)
    t_len = torch.zeros(n, dtype=torch.int64, device=dev)
    t_st = torch.zeros(n, dtype=torch.int32, device=dev)
    t_det = torch.zeros(n, dtype=torch.int32, device=dev)
    codec = lzw_b200.Codec(0)
    for name, ptr in (("device memory", d_in.data_ptr()), ("pinned host memory, in place", h_in.data_ptr()),
                      ("device memory", d_in.data_ptr())):
        ts = []
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            codec.encode_batch_device(tiff_params(), n, ptr, t_off.data_ptr(), t_out.data_ptr(), t_slots.data_ptr(),
                                      t_len.data_ptr(), t_st.data_ptr(), t_det.data_ptr(),
                                      stream=torch.cuda.current_stream().cuda_stream)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(f"encode launch, input in {name}: {min(ts[1:]):.2f} ms ({buf.size / min(ts[1:]) / 1e6:.1f} GB/s)", flush=True)
    # the same launch while the copy engines move pinned memory in one or both directions
    h_a = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
    h_b = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
    d_b = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    for name, h2d, d2h in (("H2D copies running", True, False), ("D2H copies running", False, True),
                           ("both running", True, True), ("no copies", False, False)):
        ts = []
        for _ in range(3):
            torch.cuda.synchronize()
            for _ in range(8):  # 8 GiB per direction: longer than the launch
                if h2d:
                    with torch.cuda.stream(s1):
                        d_a.copy_(h_a, non_blocking=True)
                if d2h:
                    with torch.cuda.stream(s2):
                        h_b.copy_(d_b, non_blocking=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            codec.encode_batch_device(tiff_params(), n, d_in.data_ptr(), t_off.data_ptr(), t_out.data_ptr(),
                                      t_slots.data_ptr(), t_len.data_ptr(), t_st.data_ptr(), t_det.data_ptr(),
                                      stream=torch.cuda.current_stream().cuda_stream)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(f"encode launch, input in device memory, {name}: {min(ts):.2f} ms", flush=True)
    codec.close()


if __name__ == "__main__":
    main()
