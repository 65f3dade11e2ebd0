import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
import torch
import lzw_b200
from lzw_b200 import workloads as W
from lzw_b200.types import tiff_params
n = 65536
buf, off = W.tiff_strips(n)
dev = torch.device("cuda:0")
d_in = torch.from_numpy(buf).to(dev)
t_off = torch.from_numpy(off.view(np.int64)).to(dev)
codec = lzw_b200.Codec(0)
total = int(off[-1])
for lo_f, hi_f in ((0.0, 0.103), (0.103, 0.515), (0.515, 0.762), (0.762, 0.91), (0.91, 1.0), (0.0, 1.0)):
    lo = int(np.searchsorted(off, np.uint64(total * lo_f)))
    hi = int(np.searchsorted(off, np.uint64(total * hi_f))) if hi_f < 1 else n
    m = hi - lo
    sub = off[lo:hi + 1]
    slots = W.encode_slots(sub - sub[0])
    t_slots = torch.from_numpy(slots.view(np.int64)).to(dev)
    t_out = torch.empty(int(slots[-1]), dtype=torch.uint8, device=dev)
    t_len = torch.zeros(m, dtype=torch.int64, device=dev)
    t_st = torch.zeros(m, dtype=torch.int32, device=dev)
    t_det = torch.zeros(m, dtype=torch.int32, device=dev)
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        codec.encode_batch_device(tiff_params(), m, d_in.data_ptr(), t_off.data_ptr() + 8 * lo, t_out.data_ptr(),
                                  t_slots.data_ptr(), t_len.data_ptr(), t_st.data_ptr(), t_det.data_ptr(),
                                  stream=torch.cuda.current_stream().cuda_stream)
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    nbytes = int(sub[-1] - sub[0])
    print(f"streams [{lo}, {hi}): {m} streams, {nbytes/1e9:.2f} GB, encode {min(ts):.2f} ms = {nbytes/min(ts)/1e6:.1f} GB/s", flush=True)
