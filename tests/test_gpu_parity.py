"""GPU parity: the CUDA path through the C ABI vs the CPU oracle, bit-exact.

Every test here calls libslzw.so (via lzw_b200) on cuda:0 and compares out_len / status /
detail / bytes with the oracle on the same seeded inputs.
"""
import hashlib
import io

import numpy as np
import pytest

from oracle import oracle as O
from tests import cases as T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["default", "latency-variant"])
def codec(request):
    """Every test of this module runs twice: with the encoder the library picks by batch size, and
    with the latency variant forced wherever it exists (fixed flavour; SLZW_ENC_CONFIG is read by
    slzw_create).  The throughput kernel (tensor memory + ldmatrix buckets) handles every batch of
    a variable flavour in both runs, whatever its size."""
    import os
    import lzw_b200
    old = os.environ.get("SLZW_ENC_CONFIG")
    if request.param == "latency-variant":
        os.environ["SLZW_ENC_CONFIG"] = "1"
    else:
        os.environ.pop("SLZW_ENC_CONFIG", None)
    c = lzw_b200.Codec(0)
    yield c
    c.close()
    if old is None:
        os.environ.pop("SLZW_ENC_CONFIG", None)
    else:
        os.environ["SLZW_ENC_CONFIG"] = old
    lzw_b200.Codec(0).close()  # the setting is process-wide: back to what the environment says


def gp(p):
    """oracle Params -> library Params (same field layout)."""
    from lzw_b200 import _lib
    return _lib.Params(p.flavour, p.code_size, p.big_endian, p.tiff_early_change)


# ---- reference known-answer vectors, now through the GPU ---------------------------------------
def test_kat_encode(codec, lorem, lorem_encoded):
    assert codec.encode(gp(O.variable(2, False, False)), T.D40)[2] == bytes(
        [0x8C, 0x2D, 0x99, 0x87, 0x2A, 0x1C, 0xDC, 0x33, 0xA0, 0x02, 0x55, 0x00])   # encoder.rs:666-686
    assert codec.encode(gp(O.gif(2)), bytes([0, 0, 1, 3])) == (0, 0, bytes([0x04, 0x32, 0x05]))
    assert codec.encode(gp(O.tiff()), bytes([0, 0, 1, 3])) == (0, 0, bytes([0x80, 0, 0, 0, 0x10, 0x1C, 0x04]))
    assert codec.encode(gp(O.fixed()), bytes([0, 0, 1, 3])) == (0, 0, bytes([0, 0, 0, 1, 0x30, 0]))
    st, _, out = codec.encode(gp(O.gif(7)), lorem)                                    # encoder.rs:740-755
    assert st == 0 and out == lorem_encoded
    assert codec.encode(gp(O.variable(10, False, False)), bytes([0]))[:2] == (O.ERR_CODE_SIZE, 10)
    assert codec.encode(gp(O.variable(2, True, False)), bytes([0, 1, 8, 3]))[:2] == (O.ERR_UNEXPECTED_CODE, 8)


def test_kat_decode(codec, lorem, lorem_encoded):
    assert codec.decode(gp(O.gif(7)), lorem_encoded) == (0, 0, lorem)                # decoder.rs:703-718
    assert codec.decode(gp(O.variable(2, False, False)), bytes(
        [0x8C, 0x2D, 0x99, 0x87, 0x2A, 0x1C, 0xDC, 0x33, 0xA0, 0x02, 0x55, 0x00])) == (0, 0, T.D40)
    assert codec.decode(gp(O.variable(10, False, False)), bytes([0]))[:2] == (O.ERR_CODE_SIZE, 10)
    st, detail, _ = codec.decode(gp(O.tiff()), T.BAD_TIFF)                           # decoder.rs:759-769
    assert (st, detail) == (O.ERR_UNEXPECTED_CODE, 258)


def test_facade_matches_reference_api(lorem, lorem_encoded):
    """The reference-shaped facade (same names / arguments / error text)."""
    from lzw_b200 import CodeSizeStrategy, Endianness
    from lzw_b200.decoder import (DecodingError, FixedDecoder, GifStyleDecoder, TiffStyleDecoder,
                                  VariableDecoder)
    from lzw_b200.encoder import (EncodingError, FixedEncoder, GifStyleEncoder, TiffStyleEncoder,
                                  VariableEncoder)
    data = bytes([0, 0, 1, 3])
    assert GifStyleEncoder.encode_to_vec(data, 2) == bytes([0x04, 0x32, 0x05])
    assert TiffStyleEncoder.encode_to_vec(data) == bytes([0x80, 0, 0, 0, 0x10, 0x1C, 0x04])
    assert FixedEncoder.encode_to_vec(data, Endianness.LittleEndian) == bytes([0, 0, 0, 1, 0x30, 0])
    assert VariableEncoder.encode_to_vec(data, 2, Endianness.LittleEndian, CodeSizeStrategy.Default) == bytes([0x04, 0x32, 0x05])
    assert GifStyleDecoder.decode_to_vec(bytes([0x04, 0x32, 0x05]), 2) == data
    assert TiffStyleDecoder.decode_to_vec(bytes([0x80, 0, 0, 0, 0x10, 0x1C, 0x04])) == data
    assert FixedDecoder.decode_to_vec(bytes([0, 0, 0, 1, 0x30, 0]), Endianness.LittleEndian) == data
    sink = io.BytesIO()
    VariableEncoder.encode(io.BytesIO(lorem), sink, 7, Endianness.LittleEndian, CodeSizeStrategy.Default)
    assert sink.getvalue() == lorem_encoded                                           # examples/usage.rs
    out = bytearray()
    VariableDecoder.decode(lorem_encoded, out, 7, Endianness.LittleEndian, CodeSizeStrategy.Default)
    assert bytes(out) == lorem
    with pytest.raises(EncodingError) as e:                                           # encoder.rs:758-774
        VariableEncoder.encode(bytes([0]), bytearray(), 10, Endianness.LittleEndian, CodeSizeStrategy.Default)
    assert str(e.value) == "Code size must be between 2 and 8, was 10."
    with pytest.raises(EncodingError) as e:                                           # encoder.rs:777-795
        VariableEncoder.encode_to_vec(bytes([0, 1, 8, 3]), 2, Endianness.BigEndian, CodeSizeStrategy.Default)
    assert str(e.value) == "Unexpected code 8. For code size 2, data should be < 4."
    with pytest.raises(DecodingError) as e:                                           # decoder.rs:721-737
        VariableDecoder.decode(bytes([0]), bytearray(), 10, Endianness.LittleEndian, CodeSizeStrategy.Default)
    assert str(e.value) == "Code size must be between 2 and 8, was 10"
    with pytest.raises(DecodingError) as e:                                           # decoder.rs:759-769
        TiffStyleDecoder.decode_to_vec(T.BAD_TIFF)
    assert str(e.value) == "Unexpected code while decompressing: 258"


# ---- config 1: lorem through every flavour -------------------------------------------------------
@pytest.mark.parametrize("p", T.all_params(), ids=T.pname)
def test_lorem_every_flavour(codec, lorem, p):
    data = bytes(b & T.max_symbol(p) for b in lorem) if T.max_symbol(p) < 127 else lorem
    want = O.encode(p, data)
    got = codec.encode(gp(p), data)
    assert got == want
    assert codec.decode(gp(p), got[2]) == O.decode(p, want[2])


# ---- config 2: sunflower strips, TIFF ----------------------------------------------------------------
def test_sunflower_strips_tiff(codec):
    from lzw_b200 import workloads as W
    px, off = W.sunflower_strips()
    out, out_off, out_len, status, detail = codec.encode_batch(gp(O.tiff()), px, off)
    assert (status == 0).all()
    assert out_len.tolist() == [5712, 6461, 5872, 5996, 7242, 7983, 8079, 8407, 8190, 7944, 6900, 3618, 1982, 561]
    dense, doff = T.pack_dense(out, out_off, out_len)
    assert hashlib.sha256(dense.tobytes()).hexdigest().startswith("2a13fbbe2522f966")
    o_out, o_len, o_st, _ = O.encode_batch(O.tiff(), px, off, out_off)
    assert np.array_equal(o_len, out_len) and T.slots_equal(out, o_out, out_off, out_len) < 0
    dec, dlen, dst, _ = codec.decode_batch(gp(O.tiff()), dense, doff, off)
    assert (dst == 0).all() and np.array_equal(dlen, np.diff(off)) and np.array_equal(dec[:px.size], px)


# ---- seeded ragged batches, every flavour ------------------------------------------------------------
@pytest.mark.parametrize("p", T.all_params(), ids=T.pname)
def test_ragged_batch_encode_decode(codec, p):
    buf, off = T.make_batch(1234 + p.code_size * 4 + p.big_endian * 2 + p.tiff_early_change,
                            96, T.max_symbol(p))
    out, out_off, out_len, status, detail = codec.encode_batch(gp(p), buf, off)
    o_out, o_len, o_st, o_det = O.encode_batch(p, buf, off, out_off)
    assert np.array_equal(status, o_st) and np.array_equal(detail, o_det)
    assert np.array_equal(out_len, o_len)
    assert T.slots_equal(out, o_out, out_off, out_len) < 0
    dense, doff = T.pack_dense(out, out_off, out_len)
    # decode into slots two bytes larger than needed
    cap = np.zeros(off.size, dtype=np.uint64)
    cap[1:] = np.cumsum(np.diff(off) + np.uint64(2))
    dec, dlen, dst, ddet = codec.decode_batch(gp(p), dense, doff, cap)
    o_dec, o_dlen, o_dst, o_ddet = O.decode_batch(p, dense, doff, cap)
    assert np.array_equal(dst, o_dst) and np.array_equal(ddet, o_ddet) and np.array_equal(dlen, o_dlen)
    assert T.slots_equal(dec, o_dec, cap, dlen) < 0
    # round trip wherever the reference's own decoder accepts its encoder's output
    for i in np.nonzero(o_dst == 0)[0]:
        a, b = int(cap[i]), int(off[i])
        l = int(off[i + 1] - off[i])
        assert np.array_equal(dec[a:a + l], buf[b:b + l])


def test_long_streams_with_dictionary_resets(codec):
    rng = np.random.default_rng(7)
    for p, n in [(O.tiff(), 200_000), (O.gif(8), 150_000), (O.gif(2), 120_000), (O.fixed(True), 100_000),
                 (O.variable(5, True, False), 90_000)]:
        for kind in ("random", "walk", "runs", "zeros"):
            data = T.make_stream(rng, kind, n, T.max_symbol(p)).tobytes()
            want = O.encode(p, data)
            got = codec.encode(gp(p), data)
            assert got == want, (T.pname(p), kind)
            assert codec.decode(gp(p), got[2]) == O.decode(p, want[2]), (T.pname(p), kind)


# ---- edge cases of the encoder ------------------------------------------------------------------------
@pytest.mark.parametrize("p", T.all_params(), ids=T.pname)
def test_encoder_edges(codec, p):
    hi = T.max_symbol(p)
    for data in (b"", bytes([0]), bytes([hi]), bytes([hi, hi]), bytes([0, hi, 0, hi, hi, hi, hi])):
        assert codec.encode(gp(p), data) == O.encode(p, data)
    if p.flavour == 0 and p.code_size < 8:
        # unchecked first byte (encoder.rs:311): treated as a node, masked, or a reference panic
        clear = 1 << p.code_size
        for data in (bytes([200]), bytes([200, 1]), bytes([200, 255]), bytes([clear, 1, 1, clear & hi]),
                     bytes([clear + 1, 0, 0, 0, 0]), bytes([clear, clear]), bytes([0, 0, hi + 1]),
                     bytes([clear + 2, 0]), bytes([clear + 2])):
            assert codec.encode(gp(p), data) == O.encode(p, data), data


@pytest.mark.parametrize("p", [O.gif(8), O.tiff(), O.gif(3), O.fixed(False), O.fixed(True)], ids=T.pname)
def test_encoder_output_slot_too_small(codec, p):
    """`&mut [u8]` writer semantics: fill the slot, then Io(WriteZero)."""
    rng = np.random.default_rng(5)
    data = T.make_stream(rng, "walk", 3000, T.max_symbol(p)).tobytes()
    full = O.encode(p, data)[2]
    for cap in [0, 1, 2, 3, 7, 100, len(full) - 2, len(full) - 1, len(full), len(full) + 1]:
        assert codec.encode(gp(p), data, cap=cap) == O.encode(p, data, cap=cap), cap
    # an input error that comes before / after the slot fills up
    if p.flavour == 0 and p.code_size < 8:
        bad = bytearray(data)
        bad[1500] = 255
        for cap in [10, 600, 5000]:
            assert codec.encode(gp(p), bytes(bad), cap=cap) == O.encode(p, bytes(bad), cap=cap), cap


# ---- decoder error parity --------------------------------------------------------------------------------
def _mutations(rng, packed: bytes, count: int):
    yield packed
    for cut in sorted(set(int(x) for x in rng.integers(0, len(packed), size=count))):
        yield packed[:cut]                                   # truncated: Io(UnexpectedEof) or worse
    for _ in range(count):
        b = bytearray(packed)
        for _ in range(int(rng.integers(1, 4))):
            b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
        yield bytes(b)                                       # bit flips: any decoder error, stale-table paths
    for _ in range(count // 2):
        yield bytes(rng.integers(0, 256, size=int(rng.integers(1, 400)), dtype=np.uint8))  # garbage


@pytest.mark.parametrize("p", [O.gif(8), O.tiff(), O.gif(2), O.gif(5), O.variable(4, True, True),
                               O.fixed(False), O.fixed(True)], ids=T.pname)
def test_decoder_error_parity_on_corrupt_streams(codec, p):
    rng = np.random.default_rng(99 + p.code_size)
    hi = T.max_symbol(p)
    streams = []
    for kind, n in (("text", 5000), ("runs", 9000), ("random", 2500)):
        packed = O.encode(p, T.make_stream(rng, kind, n, hi).tobytes())[2]
        streams += list(_mutations(rng, packed, 24))
    off = np.zeros(len(streams) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(s) for s in streams])
    buf = np.frombuffer(b"".join(streams), dtype=np.uint8)
    # generous slots: corrupt streams may expand a lot
    cap = (np.arange(len(streams) + 1, dtype=np.uint64)) * np.uint64(1 << 17)
    dec, dlen, dst, ddet = codec.decode_batch(gp(p), buf, off, cap)
    o_dec, o_dlen, o_dst, o_ddet = O.decode_batch(p, buf, off, cap)
    assert np.array_equal(dst, o_dst)
    assert np.array_equal(ddet, o_ddet)
    assert np.array_equal(dlen, o_dlen)
    assert T.slots_equal(dec, o_dec, cap, dlen) < 0
    assert len(set(dst.tolist())) >= (3 if p.flavour == 0 else 2)   # several outcomes are exercised


def test_decoder_missing_clear_code(codec):
    """decoder.rs:281-283: a stream that keeps going after the table is full."""
    # 12-bit codes for 'a' forever, LSB first, cs=8: clear, then 4000 literal codes
    codes = [256] + [97] * 4000
    widths = []
    nxt, w = 258, 9
    for i, _ in enumerate(codes):
        widths.append(w)
        if i >= 2:
            nxt += 1
            if nxt == (1 << w) and w < 12:
                w += 1
    packed = O.bitwrite(False, codes, widths)
    want = O.decode(O.gif(8), packed)
    assert want[0] == O.ERR_MISSING_CLEAR_CODE
    assert codec.decode(gp(O.gif(8)), packed) == want


def test_decoder_self_inconsistent_streams(codec):
    """encoder output the reference's own decoder rejects; parity is on the result."""
    rng = np.random.default_rng(11)
    seen = set()
    for p in (O.tiff(), O.gif(8), O.gif(3), O.variable(8, True, False), O.variable(6, False, True)):
        # final dictionary size == 2^w - inc  <=>  n such that (first + n - 1) hits the mask
        for w in range(p.code_size + 1, 12):
            n = (1 << w) - (1 if p.tiff_early_change else 0) - ((1 << p.code_size) + 2) + 1
            for _ in range(3):
                # distinct-pair data so that every byte after the first adds an entry
                data = T.make_stream(rng, "random", n, T.max_symbol(p)).tobytes()
                packed = O.encode(p, data)[2]
                want = O.decode(p, packed)
                seen.add(want[0])
                assert codec.decode(gp(p), packed) == want
    assert len(seen) >= 2


def test_decoder_output_slot_too_small(codec):
    rng = np.random.default_rng(3)
    for p in (O.tiff(), O.gif(4), O.fixed(False)):
        data = T.make_stream(rng, "runs", 4000, T.max_symbol(p)).tobytes()
        packed = O.encode(p, data)[2]
        for cap in (0, 1, 2, 17, 1000, 3999, 4000, 4001):
            assert codec.decode(gp(p), packed, cap=cap) == O.decode(p, packed, cap=cap), cap


def test_decoded_size_prepass(codec, lorem_encoded):
    assert codec.decoded_size(gp(O.gif(7)), lorem_encoded) == (0, 0, 23336)
    assert codec.decoded_size(gp(O.tiff()), T.BAD_TIFF) == O.decoded_size(O.tiff(), T.BAD_TIFF)


# ---- batch plumbing ------------------------------------------------------------------------------------------
def test_per_stream_code_size(codec):
    rng = np.random.default_rng(21)
    cs = (2 + np.arange(40) % 7).astype(np.uint8)
    streams = [T.make_stream(rng, T.KINDS[i % 6], int(rng.integers(0, 5000)), (1 << int(cs[i])) - 1) for i in range(40)]
    off = np.zeros(41, dtype=np.uint64)
    off[1:] = np.cumsum([s.size for s in streams])
    buf = np.concatenate(streams)
    out, out_off, out_len, status, detail = codec.encode_batch(gp(O.gif(8)), buf, off, code_size=cs)
    o_out, o_len, o_st, o_det = O.encode_batch(O.gif(8), buf, off, out_off, code_size=cs)
    assert np.array_equal(status, o_st) and np.array_equal(out_len, o_len)
    assert T.slots_equal(out, o_out, out_off, out_len) < 0
    dense, doff = T.pack_dense(out, out_off, out_len)
    dec, dlen, dst, _ = codec.decode_batch(gp(O.gif(8)), dense, doff, off, code_size=cs)
    o_dec, o_dlen, o_dst, _ = O.decode_batch(O.gif(8), dense, doff, off, code_size=cs)
    assert np.array_equal(dst, o_dst) and np.array_equal(dlen, o_dlen)
    assert T.slots_equal(dec, o_dec, off, dlen) < 0


def test_device_pointer_api_and_compaction(codec):
    """Device-resident batch through torch tensors + the compaction stage, on a side stream."""
    import torch
    from lzw_b200 import workloads as W
    buf, off = W.tiff_strips(256, seed=77)
    slots = W.encode_slots(off)
    n = off.size - 1
    dev = torch.device("cuda:0")
    t_in = torch.from_numpy(buf).to(dev)
    t_off = torch.from_numpy(off.view(np.int64)).to(dev)
    t_slots = torch.from_numpy(slots.view(np.int64)).to(dev)
    t_out = torch.zeros(int(slots[-1]), dtype=torch.uint8, device=dev)
    t_len = torch.zeros(n, dtype=torch.int64, device=dev)
    t_st = torch.zeros(n, dtype=torch.int32, device=dev)
    t_det = torch.zeros(n, dtype=torch.int32, device=dev)
    t_dense = torch.zeros(int(slots[-1]), dtype=torch.uint8, device=dev)
    t_doff = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        codec.encode_batch_device(gp(O.tiff()), n, t_in.data_ptr(), t_off.data_ptr(), t_out.data_ptr(),
                                  t_slots.data_ptr(), t_len.data_ptr(), t_st.data_ptr(), t_det.data_ptr(),
                                  stream=stream.cuda_stream)
        codec.compact_device(t_out.data_ptr(), t_slots.data_ptr(), t_len.data_ptr(), n,
                             t_dense.data_ptr(), t_doff.data_ptr(), align=1, stream=stream.cuda_stream)
    stream.synchronize()
    o_out, o_len, o_st, _ = O.encode_batch(O.tiff(), buf, off, slots, threads=4)
    assert np.array_equal(t_len.cpu().numpy().astype(np.uint64), o_len)
    assert (t_st.cpu().numpy() == 0).all()
    dense, doff = T.pack_dense(o_out, slots, o_len)
    assert np.array_equal(t_doff.cpu().numpy().astype(np.uint64), doff)
    assert np.array_equal(t_dense.cpu().numpy()[:dense.size], dense)
    # decode the dense buffer back on the device
    t_dec = torch.zeros(buf.size, dtype=torch.uint8, device=dev)
    with torch.cuda.stream(stream):
        codec.decode_batch_device(gp(O.tiff()), n, t_dense.data_ptr(), t_doff.data_ptr(), t_dec.data_ptr(),
                                  t_off.data_ptr(), t_len.data_ptr(), t_st.data_ptr(), t_det.data_ptr(),
                                  stream=stream.cuda_stream)
    stream.synchronize()
    st = t_st.cpu().numpy()
    o_dec, o_dlen, o_dst, _ = O.decode_batch(O.tiff(), dense, doff, off, threads=4)
    assert np.array_equal(st.astype(np.uint32), o_dst)
    ok = o_dst == 0
    got = t_dec.cpu().numpy()
    for i in np.nonzero(ok)[0]:
        assert np.array_equal(got[int(off[i]):int(off[i + 1])], buf[int(off[i]):int(off[i + 1])])


def test_encode_batch_dense_host_api(codec):
    """slzw_encode_batch_host_dense: strips back to back, offsets returned, only encoded bytes copied."""
    buf, off = T.make_batch(4242, 300, 255, max_len=20000)
    for align in (1, 2):
        dense, doff, st, det = codec.encode_batch_dense(gp(O.tiff()), buf, off, align=align)
        slots = np.zeros(off.size, dtype=np.uint64)
        slots[1:] = np.cumsum(np.diff(off) * np.uint64(2) + np.uint64(16))
        o_out, o_len, o_st, o_det = O.encode_batch(O.tiff(), buf, off, slots)
        assert np.array_equal(st, o_st) and np.array_equal(det, o_det)
        padded = (o_len + np.uint64(align - 1)) // np.uint64(align) * np.uint64(align)
        assert np.array_equal(np.diff(doff), padded)
        for i in range(o_len.size):
            assert np.array_equal(dense[int(doff[i]):int(doff[i]) + int(o_len[i])],
                                  o_out[int(slots[i]):int(slots[i]) + int(o_len[i])]), i


def test_compaction_alignment(codec):
    import torch
    rng = np.random.default_rng(5)
    n = 300
    lens = rng.integers(0, 900, size=n).astype(np.uint64)
    src_off = np.zeros(n + 1, dtype=np.uint64)
    src_off[1:] = np.cumsum(lens + rng.integers(0, 7, size=n).astype(np.uint64))
    src = rng.integers(0, 256, size=int(src_off[-1]) + 8, dtype=np.uint8)
    dev = torch.device("cuda:0")
    for align in (1, 2, 16):
        t_src = torch.from_numpy(src).to(dev)
        t_so = torch.from_numpy(src_off.view(np.int64)).to(dev)
        t_len = torch.from_numpy(lens.view(np.int64)).to(dev)
        t_dst = torch.zeros(int(src_off[-1]) + 16 * n, dtype=torch.uint8, device=dev)
        t_do = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        codec.compact_device(t_src.data_ptr(), t_so.data_ptr(), t_len.data_ptr(), n, t_dst.data_ptr(),
                             t_do.data_ptr(), align=align)
        torch.cuda.synchronize()
        do = t_do.cpu().numpy().astype(np.uint64)
        padded = (lens + np.uint64(align - 1)) // np.uint64(align) * np.uint64(align)
        want = np.zeros(n + 1, dtype=np.uint64)
        want[1:] = np.cumsum(padded)
        assert np.array_equal(do, want)
        dst = t_dst.cpu().numpy()
        for i in range(n):
            assert np.array_equal(dst[int(do[i]):int(do[i]) + int(lens[i])],
                                  src[int(src_off[i]):int(src_off[i]) + int(lens[i])]), (align, i)


def test_config3_scaled_down_vs_oracle(codec):
    """BASELINE config 3 at 2,048 strips: byte-exact against the oracle, then decode."""
    from lzw_b200 import workloads as W
    buf, off = W.tiff_strips(2048)
    out, out_off, out_len, status, detail = codec.encode_batch(gp(O.tiff()), buf, off)
    o_out, o_len, o_st, _ = O.encode_batch(O.tiff(), buf, off, out_off, threads=8)
    assert np.array_equal(status, o_st) and np.array_equal(out_len, o_len)
    assert T.slots_equal(out, o_out, out_off, out_len) < 0
    dense, doff = T.pack_dense(out, out_off, out_len)
    dec, dlen, dst, ddet = codec.decode_batch(gp(O.tiff()), dense, doff, off)
    o_dec, o_dlen, o_dst, o_ddet = O.decode_batch(O.tiff(), dense, doff, off, threads=8)
    assert np.array_equal(dst, o_dst) and np.array_equal(ddet, o_ddet) and np.array_equal(dlen, o_dlen)
    assert T.slots_equal(dec, o_dec, off, dlen) < 0
    ok = dst == 0
    assert ok.mean() > 0.95
    for i in np.nonzero(ok)[0]:
        assert np.array_equal(dec[int(off[i]):int(off[i + 1])], buf[int(off[i]):int(off[i + 1])])


# ---- host pipeline (chunked H2D / kernels / D2H) -----------------------------------------------
@pytest.mark.parametrize("dense_encode", ["streaming", "chunked"])
def test_host_paths_with_many_small_chunks(monkeypatch, dense_encode):
    """The host entry points split a batch into chunks of streams and pipeline them over several
    CUDA streams; with a tiny chunk size every code path of those pipelines runs on a small batch.
    The dense encode runs once as ONE streaming launch fed window by window (the default) and once
    through the chunked pipeline (what calls with the predictor, two-phase calls and very large
    calls take)."""
    import lzw_b200
    monkeypatch.setenv("SLZW_HOST_CHUNK_BYTES", "20000")
    monkeypatch.setenv("SLZW_HOST_ENC_STREAM", "1" if dense_encode == "streaming" else "0")
    c = lzw_b200.Codec(0)
    try:
        for p in (O.tiff(), O.gif(5), O.fixed(True)):
            buf, off = T.make_batch(99, 200, T.max_symbol(p), max_len=9000)
            slots = np.zeros(off.size, dtype=np.uint64)
            slots[1:] = np.cumsum([O.encode_bound(int(l)) + 3 for l in np.diff(off)])
            out, _, out_len, st, det = c.encode_batch(gp(p), buf, off, out_off=slots)
            o_out, o_len, o_st, o_det = O.encode_batch(p, buf, off, slots)
            assert np.array_equal(out_len, o_len) and np.array_equal(st, o_st) and np.array_equal(det, o_det)
            assert T.slots_equal(out, o_out, slots, o_len) == -1
            dense, doff, st2, det2 = c.encode_batch_dense(gp(p), buf, off, align=2)
            assert np.array_equal(st2, o_st)
            for i in range(o_len.size):
                assert np.array_equal(dense[int(doff[i]):int(doff[i]) + int(o_len[i])],
                                      o_out[int(slots[i]):int(slots[i]) + int(o_len[i])]), i
            d_in, d_off = T.pack_dense(o_out, slots, o_len)
            dec, dlen, dst, ddet = c.decode_batch(gp(p), d_in, d_off, off)
            o_dec, o_dlen, o_dst, o_ddet = O.decode_batch(p, d_in, d_off, off)
            assert np.array_equal(dlen, o_dlen) and np.array_equal(dst, o_dst) and np.array_equal(ddet, o_ddet)
            assert T.slots_equal(dec, o_dec, off, o_dlen) == -1
    finally:
        c.close()


@pytest.mark.parametrize("window_mb, reserved", [("0.03", "4"), ("0.2", "0"), ("64", "8")])
def test_streaming_dense_encode_windows_errors_and_capacity(monkeypatch, window_mb, reserved):
    """slzw_encode_batch_host_dense as one streaming launch: many windows (down to a handful of
    streams each), rejected bytes, empty streams, per-stream code sizes, and a dense buffer that is
    too small (SLZW_RC_NOMEM with the size needed, then the same call again with room)."""
    import lzw_b200
    monkeypatch.setenv("SLZW_HOST_WINDOW_MB", window_mb)
    monkeypatch.setenv("SLZW_HOST_STREAM_SMS", reserved)
    c = lzw_b200.Codec(0)
    try:
        rng = np.random.default_rng(2024)
        n = 2500
        cs = rng.integers(2, 9, size=n).astype(np.uint8)
        streams = []
        for i in range(n):
            hi = (1 << int(cs[i])) - 1
            length = int(rng.choice([0, 1, 3, 60, 900, 5000, 30000], p=[.03, .02, .02, .2, .43, .25, .05]))
            s = T.make_stream(rng, T.KINDS[i % len(T.KINDS)], length, hi)
            if hi < 255 and i % 89 == 7 and s.size > 4:
                s[int(rng.integers(1, s.size))] = hi + 1  # UnexpectedCode
            streams.append(s)
        off = np.zeros(n + 1, dtype=np.uint64)
        off[1:] = np.cumsum([s.size for s in streams])
        buf = np.concatenate(streams)
        p = O.gif(8)
        slots = np.zeros(n + 1, dtype=np.uint64)
        slots[1:] = np.cumsum([O.encode_bound(int(s.size)) for s in streams])
        o_out, o_len, o_st, o_det = O.encode_batch(p, buf, off, slots, code_size=cs, threads=8)
        assert int((o_st != 0).sum()) > 5
        total = int(o_len.sum())
        small = np.empty(total // 2, dtype=np.uint8)
        with pytest.raises(RuntimeError):
            c.encode_batch_dense(gp(p), buf, off, code_size=cs, out=small)
        for _ in range(2):  # the second call reuses every buffer of the first
            dense, doff, st, det = c.encode_batch_dense(gp(p), buf, off, code_size=cs)
            assert np.array_equal(st, o_st) and np.array_equal(det, o_det)
            assert np.array_equal(np.diff(doff), o_len) and dense.size == total
            for i in range(n):
                assert np.array_equal(dense[int(doff[i]):int(doff[i + 1])],
                                      o_out[int(slots[i]):int(slots[i]) + int(o_len[i])]), i
    finally:
        c.close()


def test_host_paths_with_pinned_buffers_read_in_place(monkeypatch):
    """Pinned input is read by the encoder in place (no staging copy), also at odd offsets inside the
    pinned allocation and with a last stream that ends on the allocation's last byte."""
    import lzw_b200
    monkeypatch.setenv("SLZW_HOST_CHUNK_BYTES", "50000")
    c = lzw_b200.Codec(0)
    try:
        p = O.tiff()
        buf, off = T.make_batch(123, 150, 255, max_len=7000)
        for shift in (0, 1, 3):
            pin = lzw_b200.PinnedBuffer(buf.size + shift)
            view = pin.array[shift:]
            view[:] = buf
            slots = np.zeros(off.size, dtype=np.uint64)
            slots[1:] = np.cumsum([O.encode_bound(int(l)) for l in np.diff(off)])
            o_out, o_len, o_st, o_det = O.encode_batch(p, buf, off, slots)
            dense, doff, st, det = c.encode_batch_dense(gp(p), view, off)
            assert np.array_equal(st, o_st) and np.array_equal(np.diff(doff), o_len)
            for i in range(o_len.size):
                assert np.array_equal(dense[int(doff[i]):int(doff[i + 1])],
                                      o_out[int(slots[i]):int(slots[i]) + int(o_len[i])]), (shift, i)
            out, _, out_len, st2, _ = c.encode_batch(gp(p), view, off, out_off=slots)
            assert np.array_equal(out_len, o_len) and T.slots_equal(out, o_out, slots, o_len) == -1
            pin.free()
    finally:
        c.close()


# ---- fast decoder: deferral to the exact kernel ---------------------------------------------------
def test_fast_decoder_defers_what_it_cannot_reproduce(codec):
    """More than 1 MiB of output between two clear codes (20-bit offsets), a first code that is
    not a root (stale-table semantics, decoder.rs:230-236) and a missing clear code go to the exact
    kernel; the result still equals the oracle's, and the diagnostics name the deferred streams."""
    rng = np.random.default_rng(5)
    p = O.tiff()
    big = bytes(2_500_000)                                                # one dictionary generation
    st, _, big_packed = O.encode(p, big)
    assert st == 0
    small = T.make_stream(rng, "text", 5000, 255).tobytes()
    _, _, small_packed = O.encode(p, small)
    streams = [big_packed, small_packed, T.BAD_TIFF]
    off = np.zeros(len(streams) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(s) for s in streams])
    data = np.frombuffer(b"".join(streams), dtype=np.uint8)
    caps = np.zeros(len(streams) + 1, dtype=np.uint64)
    caps[1:] = np.cumsum([len(big) + 64, len(small) + 64, 4096])
    dec, dlen, dst, ddet = codec.decode_batch(gp(p), data, off, caps)
    o_dec, o_dlen, o_dst, o_ddet = O.decode_batch(p, data, off, caps)
    assert np.array_equal(dlen, o_dlen) and np.array_equal(dst, o_dst) and np.array_equal(ddet, o_ddet)
    assert T.slots_equal(dec, o_dec, caps, o_dlen) == -1
    assert int(o_dst[2]) == O.ERR_UNEXPECTED_CODE and int(o_ddet[2]) == 258   # decoder.rs:759-769
    deferred = set(int(i) for i in codec.last_deferred())
    assert 0 in deferred and 1 not in deferred


def test_fast_decoder_large_outputs_stay_on_the_fast_path(codec):
    """Outputs larger than 1 MiB are fine as long as no dictionary generation spans 1 MiB."""
    rng = np.random.default_rng(6)
    for p, kind in ((O.tiff(), "runs"), (O.gif(6), "walk"), (O.gif(8), "text")):
        raw = T.make_stream(rng, kind, 3_000_000, T.max_symbol(p)).tobytes()
        st, _, packed = O.encode(p, raw)
        assert st == 0
        got = codec.decode(gp(p), packed, cap=len(raw))
        want = O.decode(p, packed, cap=len(raw))
        assert got[0] == want[0] and got[1] == want[1] and got[2] == want[2], T.pname(p)
        assert got[2] == raw or want[0] != 0
        assert len(codec.last_deferred()) == 0, T.pname(p)


def test_fast_decoder_long_words_and_in_step_sources(codec):
    """Highly repetitive streams: words longer than the 1 KiB step window (periodic copy), chains
    of codes that name entries created in the same step (KwKwK, decoder.rs:244-250)."""
    cases = []
    for p in (O.tiff(), O.gif(2), O.gif(8), O.fixed(False), O.variable(3, True, False)):
        hi = T.max_symbol(p)
        cases.append((p, bytes(900_000)))
        cases.append((p, bytes([1 % (hi + 1), 2 % (hi + 1)]) * 300_000))
        cases.append((p, bytes([3 % (hi + 1)]) * 17 + bytes([0]) * 70_000 + bytes([1]) * 70_000))
    for p, raw in cases:
        st, _, packed = O.encode(p, raw)
        assert st == 0
        got = codec.decode(gp(p), packed, cap=len(raw))
        want = O.decode(p, packed, cap=len(raw))
        assert got[0] == want[0] and got[1] == want[1] and got[2] == want[2], T.pname(p)
        assert len(codec.last_deferred()) == 0 or want[0] != 0


# ---- BASELINE configs 4 and 5 at (scaled) full stream sizes ----------------------------------------
def _roundtrip_properties(codec, p, buf, off, code_size=None, oracle_streams=8):
    """Size-independent properties at full stream size: every stream encodes OK, the decoder
    returns exactly the input (SURVEY F1 streams: right bytes, error status allowed), compressed
    sizes and bytes equal the oracle's on a sample of streams."""
    out, out_off, out_len, st, det = codec.encode_batch(gp(p), buf, off, code_size=code_size)
    assert (st == 0).all()
    dense, doff = T.pack_dense(out, out_off, out_len)
    dec, dlen, dst, ddet = codec.decode_batch(gp(p), dense, doff, off, code_size=code_size)
    assert np.array_equal(dlen, np.diff(off))
    assert np.array_equal(dec[: int(off[-1])], buf)
    assert (dst == 0).mean() > 0.9 and set(np.unique(dst)) <= {0, O.ERR_IO_UNEXPECTED_EOF, O.ERR_UNEXPECTED_CODE}
    m = min(oracle_streams, off.size - 1)
    idx = np.linspace(0, off.size - 2, m).astype(int)
    for i in idx:
        q = O.Params(p.flavour, int(code_size[i]) if code_size is not None else p.code_size, p.big_endian,
                     p.tiff_early_change)
        raw = buf[int(off[i]):int(off[i + 1])].tobytes()
        o_st, _, o_packed = O.encode(q, raw)
        assert o_st == 0
        assert dense[int(doff[i]):int(doff[i + 1])].tobytes() == o_packed, i
        o_dst, o_ddet, o_raw = O.decode(q, o_packed, cap=len(raw))
        assert (int(dst[i]), int(ddet[i])) == (o_dst, o_ddet) and o_raw == raw
    return int(doff[-1])


def test_config4_gif_frames_full_size(codec):
    """Config 4 (scaled to 28 frames): 1024x1024 8-bit-palette GIF frames, code size 2..8 per
    stream in one batch; 1 MiB outputs sit exactly at the fast decoder's 20-bit offset limit."""
    from lzw_b200 import workloads as W
    buf, off, cs = W.gif_frames(28)
    comp = _roundtrip_properties(codec, O.gif(8), buf, off, code_size=cs, oracle_streams=7)
    assert 0 < comp < buf.size
    assert len(codec.last_deferred()) == 0


def test_config5_fixed_text_chunks_full_size(codec):
    """Config 5 (scaled to 1,024 chunks): lorem-like text in 64 KiB chunks, fixed 12-bit codes,
    both bit orders; the dictionary freezes at 4096 entries (encoder.rs:645-647)."""
    from lzw_b200 import workloads as W
    buf, off = W.text_chunks(1024, corpus_bytes=4 << 20)
    for be in (False, True):
        comp = _roundtrip_properties(codec, O.fixed(be), buf, off, oracle_streams=6)
        assert 0 < comp < buf.size
        assert len(codec.last_deferred()) == 0


# ---- batches larger than one wave of the latency variant (12 streams per SM) ---------------------
@pytest.mark.parametrize("p", [O.tiff(), O.gif(5), O.fixed(False), O.fixed(True)], ids=T.pname)
def test_large_batch_with_error_streams(codec, p):
    """More than 1,776 streams: the throughput kernel for every flavour (fixed included), with
    rejected bytes, too-small slots and streams long enough to fill and freeze / reset the table."""
    rng = np.random.default_rng(77)
    n = 2000
    hi = T.max_symbol(p)
    streams = []
    for i in range(n):
        length = int(rng.choice([0, 1, 2, 40, 700, 2500, 7000], p=[.02, .02, .02, .2, .44, .2, .1]))
        s = T.make_stream(rng, T.KINDS[i % len(T.KINDS)], length, hi)
        if p.flavour == 0 and hi < 255 and i % 97 == 5 and s.size > 4:
            s[int(rng.integers(0, s.size))] = hi + 1  # UnexpectedCode (or the unchecked first byte)
        streams.append(s)
    off = np.zeros(n + 1, dtype=np.uint64)
    off[1:] = np.cumsum([s.size for s in streams])
    buf = np.concatenate(streams)
    caps = [O.encode_bound(int(s.size)) if i % 53 else int(rng.integers(0, 30)) for i, s in enumerate(streams)]
    slots = np.zeros(n + 1, dtype=np.uint64)
    slots[1:] = np.cumsum(caps)
    out, _, out_len, st, det = codec.encode_batch(gp(p), buf, off, out_off=slots)
    o_out, o_len, o_st, o_det = O.encode_batch(p, buf, off, slots, threads=8)
    assert np.array_equal(st, o_st) and np.array_equal(det, o_det) and np.array_equal(out_len, o_len)
    assert int((o_st != 0).sum()) > 10
    for i in range(n):
        a, l = int(slots[i]), int(o_len[i])
        assert np.array_equal(out[a:a + l], o_out[a:a + l]), f"stream {i}"
