"""Container framing: GIF image data sub-blocks and TIFF strip tables around the
codec's raw LZW streams, checked with an independent decoder (PIL: its own GIF LZW decoder, libtiff
for TIFF).  CPU tests use the oracle's streams (the checker), GPU tests the codec's."""
import io
import os

import numpy as np
import pytest
from PIL import Image

from lzw_b200 import containers as K
from oracle import oracle as O
from tests import cases as T
from tests.conftest import GOLDEN


def _tokyo():
    im = Image.open(os.path.join(GOLDEN, "tokyo_128_colors.png"))   # 1024 x 684, 128 colours
    pal = im.getpalette()
    palette = [tuple(pal[3 * i:3 * i + 3]) for i in range(128)]
    return np.asarray(im).copy(), palette


def _sunflower():
    with open(os.path.join(GOLDEN, "sunflower.bmp"), "rb") as f:
        raw = np.frombuffer(f.read()[54:], dtype=np.uint8)
    rows = raw.reshape(200, 544)[:, :543].reshape(200, 181, 3)        # drop the BMP row padding
    return np.ascontiguousarray(rows[::-1, :, ::-1])                    # bottom-up BGR -> top-down RGB


def test_gif_sub_blocks_round_trip():
    rng = np.random.default_rng(1)
    for n in (0, 1, 254, 255, 256, 510, 511, 70000):
        stream = rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
        framed = K.gif_image_data(7, stream)
        assert framed[0] == 7 and framed[-1] == 0 and len(framed) == 2 + n + (n + 254) // 255
        mcs, back, pos = K.parse_gif_image_data(framed)
        assert (mcs, back, pos) == (7, stream, len(framed))


def test_oracle_gif_stream_is_a_real_gif():
    """The reference's GIF-style stream (here: the oracle's) framed as GIF image data decodes with
    PIL's own LZW decoder to the original pixels (compare_crates.rs encodes exactly this image)."""
    px, palette = _tokyo()
    st, _, stream = O.encode(O.gif(7), px.tobytes())
    assert st == 0
    data = K.write_gif(px.shape[1], px.shape[0], palette, [stream], 7)
    im = Image.open(io.BytesIO(data))
    assert im.size == (px.shape[1], px.shape[0])
    assert np.array_equal(np.asarray(im.convert("P") if im.mode != "P" else im), px)
    frames = K.read_gif_frames(data)
    assert frames == [(px.shape[1], px.shape[0], 7, stream)]


def test_oracle_tiff_strips_are_a_real_tiff():
    """TIFF-style strips inside a baseline TIFF decode with libtiff (through PIL)."""
    px = _sunflower()
    buf, off = K.strips_of_image(px, rows_per_strip=16)
    strips = []
    for i in range(off.size - 1):
        st, _, s = O.encode(O.tiff(), buf[int(off[i]):int(off[i + 1])].tobytes())
        assert st == 0
        strips.append(s)
    data = K.write_tiff_lzw(px.shape[1], px.shape[0], 3, 16, strips)
    im = Image.open(io.BytesIO(data))
    assert np.array_equal(np.asarray(im), px)
    w, h, spp, rps, back = K.read_tiff_strips(data)
    assert (w, h, spp, rps) == (181, 200, 3, 16) and back == strips


@pytest.mark.gpu
def test_gpu_gif_frames_decode_with_pil():
    """Frames encoded by the GPU codec in one batch, framed, read back by PIL; then the streams
    are pulled out of the file again and decoded by the batched GPU decoder."""
    import lzw_b200
    from lzw_b200.types import gif_params
    codec = lzw_b200.Codec(0)
    try:
        px, palette = _tokyo()
        frames = np.stack([px, px[::-1], np.roll(px, 77, axis=1)])
        off = np.arange(4, dtype=np.uint64) * np.uint64(px.size)
        dense, doff, st, _ = codec.encode_batch_dense(gif_params(7), frames.reshape(-1), off)
        assert (st == 0).all()
        data = K.write_gif(px.shape[1], px.shape[0], palette, K.split_dense(dense, doff), 7)
        im = Image.open(io.BytesIO(data))
        for i in range(3):
            im.seek(i)
            assert np.array_equal(np.asarray(im.convert("RGB")),
                                  np.asarray(palette, dtype=np.uint8)[frames[i]]), i
        got = K.read_gif_frames(data)
        streams = [g[3] for g in got]
        in_off = np.zeros(4, dtype=np.uint64)
        in_off[1:] = np.cumsum([len(s) for s in streams])
        dec, dlen, dst, _ = codec.decode_batch(gif_params(7), np.frombuffer(b"".join(streams), dtype=np.uint8),
                                               in_off, off)
        assert (dst == 0).all() and np.array_equal(dec[: frames.size], frames.reshape(-1))
    finally:
        codec.close()


@pytest.mark.gpu
def test_gpu_tiff_strips_decode_with_libtiff():
    import lzw_b200
    from lzw_b200.types import tiff_params
    codec = lzw_b200.Codec(0)
    try:
        px = _sunflower()
        for rps in (1, 16, 200):
            buf, off = K.strips_of_image(px, rows_per_strip=rps)
            dense, doff, st, _ = codec.encode_batch_dense(tiff_params(), buf, off)
            assert (st == 0).all()
            data = K.write_tiff_lzw(px.shape[1], px.shape[0], 3, rps, K.split_dense(dense, doff))
            assert np.array_equal(np.asarray(Image.open(io.BytesIO(data))), px), rps
            w, h, spp, rps2, strips = K.read_tiff_strips(data)
            in_off = np.zeros(len(strips) + 1, dtype=np.uint64)
            in_off[1:] = np.cumsum([len(s) for s in strips])
            dec, dlen, dst, _ = codec.decode_batch(tiff_params(), np.frombuffer(b"".join(strips), dtype=np.uint8),
                                                   in_off, off)
            assert np.array_equal(dec[: buf.size], buf)
    finally:
        codec.close()


# ---- streams written by other encoders ---------------------------------------------------------------
def _pil_gif(px, palette):
    im = Image.fromarray(px, mode="P")
    im.putpalette([c for rgb in palette for c in rgb])
    b = io.BytesIO()
    im.save(b, format="GIF", interlace=False)
    data = b.getvalue()
    return data, np.asarray(Image.open(io.BytesIO(data)))


def _libtiff_tiff(rgb):
    b = io.BytesIO()
    Image.fromarray(rgb).save(b, format="TIFF", compression="tiff_lzw")
    return b.getvalue()


def test_oracle_decodes_streams_of_other_encoders():
    """The reference's decoder semantics (the oracle) read what PIL's GIF encoder and libtiff's LZW
    encoder write (libtiff widens the code before the end-of-information code where salzweg's
    encoder does not -- SURVEY F1 -- which is exactly what salzweg's decoder expects)."""
    px, palette = _tokyo()
    data, want = _pil_gif(px, palette)
    (w, h, mcs, stream), = K.read_gif_frames(data)
    st, _, out = O.decode(O.gif(mcs), stream, cap=w * h)
    assert st == 0 and np.array_equal(np.frombuffer(out, dtype=np.uint8).reshape(h, w), want)
    rgb = _sunflower()
    w, h, spp, rps, strips = K.read_tiff_strips(_libtiff_tiff(rgb))
    raw, pos = rgb.tobytes(), 0
    for i, s in enumerate(strips):
        n = min(rps, h - i * rps) * w * spp
        st, _, out = O.decode(O.tiff(), s, cap=n)
        assert st == 0 and out == raw[pos:pos + n], i
        pos += n


@pytest.mark.gpu
def test_gpu_decodes_streams_of_other_encoders():
    import lzw_b200
    from lzw_b200.types import gif_params, tiff_params
    codec = lzw_b200.Codec(0)
    try:
        px, palette = _tokyo()
        data, want = _pil_gif(px, palette)
        (w, h, mcs, stream), = K.read_gif_frames(data)
        st, det, out = codec.decode(gif_params(mcs), stream, cap=w * h)
        assert (st, det) == (0, 0) and np.array_equal(np.frombuffer(out, dtype=np.uint8).reshape(h, w), want)
        rgb = _sunflower()
        w, h, spp, rps, strips = K.read_tiff_strips(_libtiff_tiff(rgb))
        in_off = np.zeros(len(strips) + 1, dtype=np.uint64)
        in_off[1:] = np.cumsum([len(s) for s in strips])
        caps = np.zeros(len(strips) + 1, dtype=np.uint64)
        caps[1:] = np.cumsum([min(rps, h - i * rps) * w * spp for i in range(len(strips))])
        dec, dlen, dst, _ = codec.decode_batch(tiff_params(), np.frombuffer(b"".join(strips), dtype=np.uint8), in_off, caps)
        assert (dst == 0).all() and dec[: rgb.size].tobytes() == rgb.tobytes()
    finally:
        codec.close()


# ---- deferred clear codes: the opt-in lenient decoder flavour ----------------------
def _gif_lzw_deferred_clear(pixels: bytes, cs: int) -> bytes:
    """A GIF LZW encoder that never sends a clear code after the first one: once the dictionary has
    4096 entries it keeps encoding with them (legal GIF, and what salzweg's own decoder rejects with
    MissingClearCode, decoder.rs:281-283).  Plain Python, test input only."""
    clear, eoi = 1 << cs, (1 << cs) + 1
    table = {}
    nxt, width = eoi + 1, cs + 1
    acc = nbits = 0
    out = bytearray()

    def put(code):
        nonlocal acc, nbits
        acc |= code << nbits
        nbits += width
        while nbits >= 8:
            out.append(acc & 0xFF)
            acc >>= 8
            nbits -= 8

    put(clear)
    prefix = pixels[0]
    for k in pixels[1:]:
        key = (prefix, k)
        if key in table:
            prefix = table[key]
            continue
        put(prefix)
        if nxt < 4096:
            table[key] = nxt
            nxt += 1
            if nxt > (1 << width) and width < 12:
                width += 1
        prefix = k
    put(prefix)
    put(eoi)
    if nbits:
        out.append(acc & 0xFF)
    return bytes(out)


def _deferred_clear_case():
    rng = np.random.default_rng(11)
    px = T.make_stream(rng, "walk", 300 * 200, 63).reshape(200, 300)     # fills the dictionary several times over
    palette = [(4 * i, 255 - 4 * i, (7 * i) % 256) for i in range(64)]
    stream = _gif_lzw_deferred_clear(px.tobytes(), 6)
    return px, palette, stream


def test_deferred_clear_stream_is_a_real_gif_and_the_oracle_modes_differ():
    px, palette, stream = _deferred_clear_case()
    im = Image.open(io.BytesIO(K.write_gif(300, 200, palette, [stream], 6)))
    assert np.array_equal(np.asarray(im), px)                            # PIL's decoder defers, too
    st, _, out = O.decode(O.gif(6), stream, cap=px.size)
    assert st == O.ERR_MISSING_CLEAR_CODE and 0 < len(out) < px.size     # the reference's behaviour
    st, det, out = O.decode(O.lenient(6), stream, cap=px.size)
    assert (st, det) == (0, 0) and out == px.tobytes()


@pytest.mark.gpu
def test_gpu_lenient_flavour_decodes_deferred_clear_streams():
    import lzw_b200
    from lzw_b200.types import CodeSizeStrategy, Endianness, gif_params, lenient_params
    codec = lzw_b200.Codec(0)
    try:
        px, palette, stream = _deferred_clear_case()
        want = O.decode(O.gif(6), stream, cap=px.size)
        assert codec.decode(gif_params(6), stream, cap=px.size) == want      # strict: MissingClearCode, same bytes
        lp = lenient_params(6, Endianness.LittleEndian, CodeSizeStrategy.Default)
        assert codec.decode(lp, stream, cap=px.size) == (0, 0, px.tobytes())
        assert len(codec.last_deferred()) == 0                               # stays on the fast kernel
        # ordinary streams decode the same in both flavours, and the exact kernel knows the flavour too
        _, _, ok_stream = O.encode(O.gif(6), px.tobytes())
        assert codec.decode(lp, ok_stream, cap=px.size) == (0, 0, px.tobytes())
        assert codec.decode(lp, stream, cap=px.size // 2) == O.decode(O.lenient(6), stream, cap=px.size // 2)
    finally:
        codec.close()
