import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
import torch, lzw_b200
from lzw_b200 import workloads as W
from lzw_b200.types import tiff_params
buf, off = W.tiff_strips(65536)
total = int(off[-1]); slots = W.encode_slots(off)
h_in = torch.from_numpy(buf).pin_memory().numpy()
h_dense = torch.empty(int(slots[-1]), dtype=torch.uint8).pin_memory().numpy()
h_dec = torch.empty(total, dtype=torch.uint8).pin_memory().numpy()
p = tiff_params()
for name, env in (("stream", "1"), ("chunked", "0")):
    os.environ["SLZW_HOST_ENC_STREAM"] = env
    c = lzw_b200.Codec(0)
    for it in range(3):
        t0 = time.perf_counter()
        doff, st, det, tot = c.encode_batch_dense_begin(p, h_in, off)
        t1 = time.perf_counter()
        c.encode_batch_dense_finish(h_dense[:tot])
        t2 = time.perf_counter()
    print(f"{name}: begin {1e3*(t1-t0):.1f} ms, finish {1e3*(t2-t1):.1f} ms", flush=True)
    c.close()
mc = lzw_b200.MultiCodec([0])
for it in range(3):
    t0 = time.perf_counter()
    dense, doff, st, det = mc.encode_batch_dense(p, h_in, off, out=h_dense)
    t1 = time.perf_counter()
    dec, dlen, dst, ddet = mc.decode_batch(p, dense, doff, off, out=h_dec)
    t2 = time.perf_counter()
print(f"multi[0]: encode {1e3*(t1-t0):.1f} ms, decode {1e3*(t2-t1):.1f} ms", flush=True)
