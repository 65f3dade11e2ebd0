"""slzw_multi_*: one host batch over the GPUs of the box, and the two-phase dense encode it is built
on, against the oracle.  Runs on one GPU as well (a MultiCodec over the same device twice shards
the batch exactly as two devices would); the test over distinct devices skips on a 1-GPU box."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import cases as T

pytestmark = pytest.mark.gpu


def _devices():
    import torch
    return torch.cuda.device_count()


def _check_against_oracle(p_oracle, buf, off, slots, out, out_len, status, detail, cs=None):
    o_out, o_len, o_st, o_det = O.encode_batch(p_oracle, buf, off, slots, code_size=cs)
    assert np.array_equal(status, o_st)
    assert np.array_equal(detail, o_det)
    assert np.array_equal(out_len, o_len)
    for i in range(o_len.size):
        a, l = int(slots[i]), int(o_len[i])
        assert np.array_equal(out[a:a + l], o_out[a:a + l]), f"stream {i}"


@pytest.mark.parametrize("devices", [[0], [0, 0], [0, 0, 0]])
def test_multi_on_one_gpu_matches_oracle(devices):
    import lzw_b200
    from lzw_b200.types import tiff_params
    from lzw_b200 import workloads as W
    buf, off = T.make_batch(21, 150, 255, max_len=9000)
    slots = W.encode_slots(off)
    mc = lzw_b200.MultiCodec(devices=devices)
    assert mc.device_count == len(devices)
    out, out_len, status, detail = mc.encode_batch(tiff_params(), buf, off, slots)
    _check_against_oracle(O.tiff(), buf, off, slots, out, out_len, status, detail)
    # dense: shards back to back, offsets absolute
    dense, doff, dst, ddet = mc.encode_batch_dense(tiff_params(), buf, off)
    assert np.array_equal(np.diff(doff), out_len) and np.array_equal(dst, status)
    for i in range(out_len.size):
        a, l = int(slots[i]), int(out_len[i])
        assert np.array_equal(dense[int(doff[i]):int(doff[i]) + l], out[a:a + l]), f"dense stream {i}"
    dec, dlen, dstat, _ = mc.decode_batch(tiff_params(), dense, doff, off)
    o_dec, o_dlen, o_dstat, _ = O.decode_batch(O.tiff(), dense, doff, off)
    assert np.array_equal(dlen, o_dlen) and np.array_equal(dstat, o_dstat)
    for i in range(dlen.size):
        a, l = int(off[i]), int(dlen[i])
        assert np.array_equal(dec[a:a + l], o_dec[a:a + l])
    mc.close()


def test_multi_with_errors_and_per_stream_code_sizes():
    import lzw_b200
    from lzw_b200.types import gif_params
    rng = np.random.default_rng(8)
    n = 90
    cs = rng.integers(2, 9, size=n).astype(np.uint8)
    parts = []
    for i in range(n):
        hi = (1 << int(cs[i])) - 1
        s = T.make_stream(rng, T.KINDS[i % len(T.KINDS)], int(rng.integers(0, 4000)), hi)
        if i % 11 == 3 and s.size > 10:
            s[s.size // 2] = min(255, hi + 1 + int(rng.integers(0, 3)))  # rejected byte (unless cs == 8)
        parts.append(s)
    off = np.zeros(n + 1, dtype=np.uint64)
    off[1:] = np.cumsum([p.size for p in parts])
    buf = np.concatenate(parts) if off[-1] else np.zeros(0, np.uint8)
    slots = np.zeros(n + 1, dtype=np.uint64)
    caps = [O.encode_bound(int(l)) if i % 7 else 5 for i, l in enumerate(np.diff(off))]  # some slots too small
    slots[1:] = np.cumsum(caps)
    mc = lzw_b200.MultiCodec(devices=[0, 0])
    out, out_len, status, detail = mc.encode_batch(gif_params(8), buf, off, slots, code_size=cs)
    _check_against_oracle(O.gif(8), buf, off, slots, out, out_len, status, detail, cs=cs)
    mc.close()


def test_two_phase_dense_encode():
    import lzw_b200
    from lzw_b200.types import tiff_params
    buf, off = T.make_batch(4, 300, 255, max_len=5000)
    codec = lzw_b200.Codec(0)
    dense, doff, st, det = codec.encode_batch_dense(tiff_params(), buf, off)
    roff, rst, rdet, total = codec.encode_batch_dense_begin(tiff_params(), buf, off)
    assert total == int(doff[-1]) and np.array_equal(roff, doff) and np.array_equal(rst, st)
    small = np.empty(max(total - 1, 1), dtype=np.uint8)
    with pytest.raises(lzw_b200.codec.SlzwError):
        codec.encode_batch_dense_finish(small)  # SLZW_RC_NOMEM, the bytes stay available
    out = codec.encode_batch_dense_finish(np.empty(total, dtype=np.uint8))
    assert np.array_equal(out, dense)
    # a second _finish has nothing left to copy; neither has one after a one-phase dense encode,
    # which reuses the device buffer the _begin left its bytes in (slzw.h)
    untouched = codec.encode_batch_dense_finish(np.full(total, 0xA5, dtype=np.uint8))
    assert np.all(untouched == 0xA5)
    codec.encode_batch_dense_begin(tiff_params(), buf, off)
    again, _, _, _ = codec.encode_batch_dense(tiff_params(), buf, off)
    assert np.array_equal(again, dense)
    untouched = codec.encode_batch_dense_finish(np.full(total, 0xA5, dtype=np.uint8))
    assert np.all(untouched == 0xA5)
    codec.close()


def test_multi_over_distinct_devices():
    if _devices() < 2:
        pytest.skip("needs at least two GPUs")
    import lzw_b200
    from lzw_b200 import workloads as W
    from lzw_b200.types import tiff_params
    buf, off = W.tiff_strips(512, seed=9)
    slots = W.encode_slots(off)
    mc = lzw_b200.MultiCodec()  # every visible device
    assert mc.device_count == _devices()
    dense, doff, st, det = mc.encode_batch_dense(tiff_params(), buf, off)
    o_out, o_len, o_st, _ = O.encode_batch(O.tiff(), buf, off, slots)
    assert np.array_equal(np.diff(doff), o_len) and np.array_equal(st, o_st)
    for i in range(o_len.size):
        a, l = int(slots[i]), int(o_len[i])
        assert np.array_equal(dense[int(doff[i]):int(doff[i]) + l], o_out[a:a + l]), f"stream {i}"
    dec, dlen, dstat, _ = mc.decode_batch(tiff_params(), dense, doff, off)
    assert np.array_equal(dec[: int(off[-1])], buf) or int((dstat != 0).sum()) > 0
    mc.close()


def test_second_concurrent_host_call_on_one_context_is_refused():
    """include/slzw.h: the *_host entry points of one context run one at a time; a call that arrives
    while another one is running returns SLZW_RC_INVALID instead of interleaving the staging slots."""
    import threading

    import lzw_b200
    from lzw_b200 import workloads as W
    from lzw_b200.types import tiff_params
    buf, off = W.tiff_strips(3000, seed=5)
    slots = W.encode_slots(off)
    codec = lzw_b200.Codec(0)
    o_out, o_len, o_st, _ = O.encode_batch(O.tiff(), buf, off, slots, threads=8)
    results, errors = [], []

    def work():
        for _ in range(6):
            try:
                out, _, out_len, st, det = codec.encode_batch(tiff_params(), buf, off, out_off=slots)
                results.append((out_len.copy(), st.copy()))
            except lzw_b200.codec.SlzwError as e:
                errors.append(str(e))

    threads = [threading.Thread(target=work) for _ in range(3)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert results, "no call went through"
    for out_len, st in results:  # every call that ran is right
        assert np.array_equal(out_len, o_len) and np.array_equal(st, o_st)
    for e in errors:  # and every refused call says why
        assert "rc=-2" in e and "in use by another host call" in e
    codec.close()
