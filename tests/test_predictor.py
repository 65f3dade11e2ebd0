"""TIFF Predictor = 2 (horizontal differencing) either side of the codec.

The reference has no predictor (it stops at the code stream), so the checker here is a numpy
restatement of TIFF 6.0 section 14 plus libtiff itself (through PIL): files written with the GPU
path must open in libtiff, files written by libtiff must decode through the GPU path."""
import io

import numpy as np
import pytest
from PIL import Image

from lzw_b200 import containers as K
from oracle import oracle as O


def hdiff_ref(strip: np.ndarray, row_bytes: int, spp: int) -> np.ndarray:
    out = strip.copy()
    for r0 in range(0, strip.size, row_bytes):
        row = strip[r0:r0 + row_bytes]
        if row.size > spp:
            out[r0 + spp:r0 + row.size] = row[spp:] - row[:-spp]        # uint8 arithmetic wraps
    return out


def hacc_ref(strip: np.ndarray, row_bytes: int, spp: int) -> np.ndarray:
    out = strip.copy()
    for r0 in range(0, strip.size, row_bytes):
        row = out[r0:r0 + row_bytes]
        for c in range(spp):
            row[c::spp] = np.cumsum(row[c::spp], dtype=np.uint64).astype(np.uint8)
    return out


def _smooth_rgb(h=96, w=160, spp=3, seed=3):
    rng = np.random.default_rng(seed)
    steps = rng.integers(-2, 3, size=(h, w, spp))
    img = (np.cumsum(steps, axis=1) + rng.integers(0, 256, size=(h, 1, spp))) % 256
    return np.ascontiguousarray(img.astype(np.uint8)[..., 0] if spp == 1 else img.astype(np.uint8))


def test_reference_restatement_round_trips_and_matches_libtiff():
    """numpy differencing + the oracle's TIFF-style encoder inside a Predictor = 2 file opens in
    libtiff; a libtiff-written Predictor = 2 file decodes with the oracle + numpy accumulation."""
    for spp in (1, 3):
        img = _smooth_rgb(spp=spp)
        row_bytes = img.shape[1] * spp
        buf, off = K.strips_of_image(img, rows_per_strip=20)
        strips = []
        for i in range(off.size - 1):
            d = hdiff_ref(buf[int(off[i]):int(off[i + 1])], row_bytes, spp)
            assert np.array_equal(hacc_ref(d, row_bytes, spp), buf[int(off[i]):int(off[i + 1])])
            st, _, s = O.encode(O.tiff(), d.tobytes())
            assert st == 0
            strips.append(s)
        data = K.write_tiff_lzw(img.shape[1], img.shape[0], spp, 20, strips, predictor=2)
        assert np.array_equal(np.asarray(Image.open(io.BytesIO(data))), img)
        info = K.read_tiff(data)
        assert info["predictor"] == 2 and info["strips"] == strips
        with pytest.raises(ValueError):
            K.read_tiff_strips(data)
        b = io.BytesIO()
        Image.fromarray(img).save(b, format="TIFF", compression="tiff_lzw", tiffinfo={317: 2})
        info = K.read_tiff(b.getvalue())
        assert info["predictor"] == 2
        pos = 0
        for i, s in enumerate(info["strips"]):
            n = min(info["rows_per_strip"], img.shape[0] - i * info["rows_per_strip"]) * row_bytes
            st, _, out = O.decode(O.tiff(), s, cap=n)
            assert st == 0
            got = hacc_ref(np.frombuffer(out, dtype=np.uint8), row_bytes, spp)
            assert got.tobytes() == img.tobytes()[pos:pos + n]
            pos += n


@pytest.fixture(scope="module")
def codec():
    import lzw_b200
    c = lzw_b200.Codec(0)
    yield c
    c.close()


def _ragged(rng, n, lo, hi):
    lens = rng.integers(lo, hi, size=n)
    gaps = rng.integers(0, 40, size=n)                      # slot layout: unused bytes between streams
    off = np.zeros(n + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens + gaps)
    return off, lens.astype(np.uint64)


@pytest.mark.gpu
@pytest.mark.parametrize("spp", [1, 2, 3, 4])
def test_gpu_predictor_kernels_match_the_restatement(codec, spp):
    """Both directions, ragged unaligned streams in a slot layout (explicit lengths), rows from one
    pixel to several warp passes, short last rows; bytes outside the streams stay untouched."""
    import torch
    rng = np.random.default_rng(100 + spp)
    dev = torch.device("cuda:0")
    for pixels, n, lo, hi in ((1, 40, 0, 300), (3, 60, 1, 2000), (50, 80, 1, 6000), (181, 64, 100, 20000),
                              (1024, 48, 1, 70000), (5000, 12, 1, 200000)):
        row_bytes = pixels * spp
        off, lens = _ragged(rng, n, lo, hi)
        base = int(rng.integers(0, 16))                     # misalign the whole batch too
        buf = rng.integers(0, 256, size=base + int(off[-1]) + 64, dtype=np.uint8)
        want_d, want_a = buf.copy(), buf.copy()
        for i in range(n):
            a, b = base + int(off[i]), base + int(off[i]) + int(lens[i])
            want_d[a:b] = hdiff_ref(buf[a:b], row_bytes, spp)
            want_a[a:b] = hacc_ref(buf[a:b], row_bytes, spp)
        t_off = torch.from_numpy(off.view(np.int64)).to(dev)
        t_len = torch.from_numpy(lens.view(np.int64)).to(dev)
        for direction, want in ((codec.DIFFERENCE, want_d), (codec.ACCUMULATE, want_a)):
            t = torch.from_numpy(buf).to(dev)
            codec.tiff_predictor_device(direction, t.data_ptr() + base, t_off.data_ptr(), n, row_bytes, spp,
                                        len_ptr=t_len.data_ptr())
            torch.cuda.synchronize()
            got = t.cpu().numpy()
            bad = np.flatnonzero(got != want)
            assert bad.size == 0, (pixels, direction, bad[:8])
        # dense layout (no length array) and the round trip difference -> accumulate
        dense_off = np.zeros(n + 1, dtype=np.uint64)
        dense_off[1:] = np.cumsum(lens)
        t = torch.from_numpy(buf).to(dev)
        t_doff = torch.from_numpy(dense_off.view(np.int64)).to(dev)
        codec.tiff_predictor_device(codec.DIFFERENCE, t.data_ptr(), t_doff.data_ptr(), n, row_bytes, spp)
        codec.tiff_predictor_device(codec.ACCUMULATE, t.data_ptr(), t_doff.data_ptr(), n, row_bytes, spp)
        torch.cuda.synchronize()
        assert np.array_equal(t.cpu().numpy(), buf), pixels


@pytest.mark.gpu
def test_gpu_predictor_rejects_what_it_does_not_cover(codec):
    from lzw_b200.codec import SlzwError
    for row_bytes, spp in ((0, 1), (10, 3), (10, 0), (10, 5)):
        with pytest.raises(SlzwError):
            codec.set_tiff_predictor(row_bytes, spp)
    codec.set_tiff_predictor(0, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("chunk", [None, "30000"])
def test_gpu_predictor_tiff_files_both_directions(monkeypatch, chunk):
    """Host pipeline with the predictor switched on: strips are differenced and encoded on the
    device (the caller's pixels stay as they were), the file opens in libtiff; a libtiff-written
    Predictor = 2 file is decoded and accumulated on the device.  With and without chunking."""
    import lzw_b200
    from lzw_b200.types import tiff_params
    if chunk:
        monkeypatch.setenv("SLZW_HOST_CHUNK_BYTES", chunk)
    codec = lzw_b200.Codec(0)
    try:
        for spp in (1, 3):
            img = _smooth_rgb(h=200, w=181, spp=spp, seed=spp)
            row_bytes = img.shape[1] * spp
            for rps in (1, 16, 200):
                buf, off = K.strips_of_image(img, rows_per_strip=rps)
                keep = buf.copy()
                codec.set_tiff_predictor(row_bytes, spp)
                dense, doff, st, _ = codec.encode_batch_dense(tiff_params(), buf, off)
                assert (st == 0).all() and np.array_equal(buf, keep)
                data = K.write_tiff_lzw(img.shape[1], img.shape[0], spp, rps, K.split_dense(dense, doff),
                                        predictor=2)
                assert np.array_equal(np.asarray(Image.open(io.BytesIO(data))), img), (spp, rps)
                # the same strips through the slot-layout entry point
                out, slots, olen, st2, _ = codec.encode_batch(tiff_params(), buf, off)
                assert (st2 == 0).all() and np.array_equal(olen, np.diff(doff))
                for i in range(off.size - 1):
                    assert out[int(slots[i]):int(slots[i]) + int(olen[i])].tobytes() == \
                        dense[int(doff[i]):int(doff[i + 1])].tobytes()
                # predictor on: the streams are the oracle's encoding of the differenced strip
                d0 = hdiff_ref(buf[: int(off[1])], row_bytes, spp)
                assert O.encode(O.tiff(), d0.tobytes())[2] == dense[: int(doff[1])].tobytes()
                # and back
                info = K.read_tiff(data)
                strips = info["strips"]
                in_off = np.zeros(len(strips) + 1, dtype=np.uint64)
                in_off[1:] = np.cumsum([len(s) for s in strips])
                dec, dlen, dst, _ = codec.decode_batch(tiff_params(), np.frombuffer(b"".join(strips), dtype=np.uint8),
                                                       in_off, off)
                assert (dst == 0).all() and np.array_equal(dec[: buf.size], buf), (spp, rps)
                codec.set_tiff_predictor(0, 0)
            # a file libtiff wrote
            b = io.BytesIO()
            Image.fromarray(img).save(b, format="TIFF", compression="tiff_lzw", tiffinfo={317: 2})
            info = K.read_tiff(b.getvalue())
            assert info["predictor"] == 2
            strips, rps = info["strips"], info["rows_per_strip"]
            in_off = np.zeros(len(strips) + 1, dtype=np.uint64)
            in_off[1:] = np.cumsum([len(s) for s in strips])
            caps = np.zeros(len(strips) + 1, dtype=np.uint64)
            caps[1:] = np.cumsum([min(rps, img.shape[0] - i * rps) * row_bytes for i in range(len(strips))])
            codec.set_tiff_predictor(row_bytes, spp)
            dec, dlen, dst, _ = codec.decode_batch(tiff_params(), np.frombuffer(b"".join(strips), dtype=np.uint8),
                                                   in_off, caps)
            codec.set_tiff_predictor(0, 0)
            assert (dst == 0).all() and dec[: img.size].tobytes() == img.tobytes()
    finally:
        codec.close()


@pytest.mark.gpu
def test_gpu_predictor_with_pinned_input_leaves_it_alone():
    """Pinned input is normally read in place by the encoder; with the predictor on it is staged,
    so the caller's pinned pixels are not differenced."""
    import lzw_b200
    from lzw_b200.types import tiff_params
    codec = lzw_b200.Codec(0)
    try:
        img = _smooth_rgb(h=64, w=300, spp=3)
        buf, off = K.strips_of_image(img, rows_per_strip=8)
        pin = lzw_b200.PinnedBuffer(buf.size)
        pin.array[:] = buf
        codec.set_tiff_predictor(900, 3)
        dense, doff, st, _ = codec.encode_batch_dense(tiff_params(), pin.array, off)
        assert (st == 0).all() and np.array_equal(pin.array, buf)
        d0 = hdiff_ref(buf[: int(off[1])], 900, 3)
        assert O.encode(O.tiff(), d0.tobytes())[2] == dense[: int(doff[1])].tobytes()
        pin.free()
    finally:
        codec.close()
