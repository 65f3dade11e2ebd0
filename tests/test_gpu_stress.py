"""Stress of the dictionary insert -> lookup ordering in the encoder's bucket step (one lane's
st.shared / tcgen05.st followed by the next step's collective load): tiny streams that do almost
nothing but inserts, many launches, every launch compared with the oracle."""
import os

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def test_insert_then_lookup_ordering_under_stress():
    import torch

    import lzw_b200
    from lzw_b200.types import tiff_params
    from lzw_b200 import workloads as W
    launches = int(os.environ.get("SLZW_STRESS_LAUNCHES", "1000"))  # 10,000: export SLZW_STRESS_LAUNCHES
    rng = np.random.default_rng(17)
    n = 4200  # a little more than the 4,144 streams one launch keeps in flight
    # random bytes: nearly every byte is a miss (insert), lengths 40-400 so that streams turn over fast
    lens = rng.integers(40, 400, size=n)
    off = np.zeros(n + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    buf = rng.integers(0, 256, size=int(off[-1]), dtype=np.uint8)
    # every 16th stream: two symbols only, so that lookups hit entries inserted a few bytes earlier
    for i in range(0, n, 16):
        a, b = int(off[i]), int(off[i + 1])
        buf[a:b] = rng.integers(0, 2, size=b - a, dtype=np.uint8) * 7
    slots = W.encode_slots(off)
    o_out, o_len, o_st, _ = O.encode_batch(O.tiff(), buf, off, slots, threads=os.cpu_count() or 1)
    dev = torch.device("cuda:0")
    t_in = torch.from_numpy(buf).to(dev)
    t_off = torch.from_numpy(off.view(np.int64)).to(dev)
    t_slots = torch.from_numpy(slots.view(np.int64)).to(dev)
    t_out = torch.zeros(int(slots[-1]), dtype=torch.uint8, device=dev)
    t_len = torch.zeros(n, dtype=torch.int64, device=dev)
    t_st = torch.zeros(n, dtype=torch.int32, device=dev)
    t_det = torch.zeros(n, dtype=torch.int32, device=dev)
    ref_len = torch.from_numpy(o_len.astype(np.int64)).to(dev)
    # expected bytes of every slot (zeros beyond the encoded length, as the zeroed output buffer)
    exp = np.zeros(int(slots[-1]), dtype=np.uint8)
    for i in range(n):
        a, l = int(slots[i]), int(o_len[i])
        exp[a:a + l] = o_out[a:a + l]
    t_exp = torch.from_numpy(exp).to(dev)
    codec = lzw_b200.Codec(0)
    os.environ["SLZW_ENC_CONFIG"] = "2"  # the throughput kernel (tensor memory + ldmatrix buckets)
    codec2 = lzw_b200.Codec(0)
    try:
        bad = 0
        for k in range(launches):
            t_out.zero_()
            codec2.encode_batch_device(tiff_params(), n, t_in.data_ptr(), t_off.data_ptr(), t_out.data_ptr(),
                                       t_slots.data_ptr(), t_len.data_ptr(), t_st.data_ptr(), t_det.data_ptr(),
                                       stream=torch.cuda.current_stream().cuda_stream)
            if not (torch.equal(t_len, ref_len) and torch.equal(t_out, t_exp) and int(t_st.abs().sum().item()) == 0):
                bad += 1
        assert bad == 0, f"{bad} of {launches} launches differ from the oracle"
    finally:
        os.environ.pop("SLZW_ENC_CONFIG", None)
        codec2.close()
        codec.close()
        lzw_b200.Codec(0).close()
