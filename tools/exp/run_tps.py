import ctypes as C, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lzw_b200 import workloads as W
lib = C.CDLL(os.path.join(os.path.dirname(__file__), "libtps.so"))
dev = torch.device("cuda:0")
nstreams = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
buf, off = W.tiff_strips(nstreams)
lens = np.diff(off).astype(np.int64)
order = np.argsort(-lens, kind="stable").astype(np.uint32)
coff = np.zeros(nstreams + 1, dtype=np.uint64); coff[1:] = np.cumsum(lens + 16)
t_in = torch.from_numpy(buf).to(dev); t_off = torch.from_numpy(off.view(np.int64)).to(dev)
t_order = torch.from_numpy(order.view(np.int32)).to(dev); t_coff = torch.from_numpy(coff.view(np.int64)).to(dev)
t_codes = torch.empty(int(coff[-1]), dtype=torch.int16, device=dev)
t_nc = torch.zeros(nstreams, dtype=torch.int32, device=dev)
t_q = torch.zeros(1, dtype=torch.int64, device=dev)
for log_slots in (13, 12):
    for blocks_per_sm, threads in ((1, 256), (2, 256), (4, 256), (8, 256), (4, 128)):
        blocks = 148 * blocks_per_sm
        nthreads = blocks * threads
        slots = 1 << log_slots
        t_tab = torch.zeros(nthreads * slots, dtype=torch.int64, device=dev)
        gen = 1
        for it in range(3):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = lib.tps_run(C.c_void_p(t_in.data_ptr()), C.c_void_p(t_off.data_ptr()), C.c_void_p(t_order.data_ptr()), nstreams,
                             C.c_void_p(t_tab.data_ptr()), C.c_void_p(t_codes.data_ptr()), C.c_void_p(t_coff.data_ptr()),
                             C.c_void_p(t_nc.data_ptr()), C.c_void_p(t_q.data_ptr()), gen, blocks, threads, log_slots, None)
            e1.record(); torch.cuda.synchronize()
            gen += 100000
            ms = e0.elapsed_time(e1)
        nc = t_nc.cpu().numpy().astype(np.int64)
        print(f"log_slots={log_slots} blocks/SM={blocks_per_sm} threads={threads} resident={nthreads} table_GB={nthreads*slots*8/1e9:.2f} "
              f"ms={ms:.2f} GB/s={buf.size/ms/1e6:.2f} codes/byte={nc.sum()/buf.size:.3f} rc={rc}", flush=True)
        del t_tab
