"""Container framing either side of the codec: GIF image data and TIFF strips.

The reference stops at the raw LZW code stream (lzw/examples/compress_image_data.rs:22-24 sinks
it).  These helpers put batches of streams produced by the GPU codec into files other software
reads, and pull the streams out of such files for the batched decoder.  They are plain host-side
byte shuffling: no LZW happens here.

GIF (GIF89a spec, "Table Based Image Data"): one byte LZW minimum code size, then the code stream
cut into sub-blocks of at most 255 bytes, each preceded by its length, then a zero-length block.
The stream is exactly what GifStyleEncoder produces (LSB-first, clear code first, EOI last,
lzw/src/encoder.rs:392-399).

TIFF (TIFF 6.0, section 13): every strip is an independent LZW stream, MSB-first with the
"early change" of the code width -- TiffStyleEncoder (lzw/src/encoder.rs:479-487); the file holds
StripOffsets / StripByteCounts / RowsPerStrip.
"""
from __future__ import annotations

import struct

import numpy as np

# ---- GIF ------------------------------------------------------------------------------------------


def gif_image_data(min_code_size: int, stream: bytes) -> bytes:
    """LZW minimum code size byte + sub-blocks (<= 255 bytes each) + block terminator."""
    if not 2 <= min_code_size <= 8:
        raise ValueError("GIF LZW minimum code size must be 2..8")
    out = bytearray([min_code_size])
    for i in range(0, len(stream), 255):
        block = stream[i:i + 255]
        out.append(len(block))
        out += block
    out.append(0)
    return bytes(out)


def parse_gif_image_data(data: bytes, pos: int = 0):
    """Inverse of gif_image_data.  Returns (min_code_size, stream, position after the terminator)."""
    min_code_size = data[pos]
    pos += 1
    stream = bytearray()
    while True:
        n = data[pos]
        pos += 1
        if n == 0:
            break
        stream += data[pos:pos + n]
        pos += n
    return min_code_size, bytes(stream), pos


def write_gif(width: int, height: int, palette, frames, min_code_size: int) -> bytes:
    """A GIF89a file with a global colour table and one image per encoded frame.

    palette: sequence of (r, g, b), at most 256 entries; frames: GIF-style LZW streams of
    width * height palette indices each, all encoded with `min_code_size`."""
    table_bits = max(1, (max(len(palette), 2) - 1).bit_length())
    colours = list(palette) + [(0, 0, 0)] * ((1 << table_bits) - len(palette))
    out = bytearray(b"GIF89a")
    out += struct.pack("<HHBBB", width, height, 0x80 | ((table_bits - 1) << 4) | (table_bits - 1), 0, 0)
    for r, g, b in colours:
        out += bytes((r, g, b))
    for stream in frames:
        out += b"\x2c" + struct.pack("<HHHHB", 0, 0, width, height, 0)
        out += gif_image_data(min_code_size, stream)
    out += b"\x3b"
    return bytes(out)


def read_gif_frames(data: bytes):
    """Every image of a GIF file as (width, height, min_code_size, LZW stream); extensions and
    colour tables are skipped.  The streams go to GifStyleDecoder / Codec.decode_batch as they are."""
    if data[:6] not in (b"GIF87a", b"GIF89a"):
        raise ValueError("not a GIF file")
    flags = data[10]
    pos = 13
    if flags & 0x80:
        pos += 3 << ((flags & 7) + 1)
    frames = []
    while pos < len(data):
        tag = data[pos]
        pos += 1
        if tag == 0x3B:  # trailer
            break
        if tag == 0x21:  # extension: label, then sub-blocks
            pos += 1
            while data[pos]:
                pos += data[pos] + 1
            pos += 1
        elif tag == 0x2C:  # image descriptor
            _, _, w, h, f = struct.unpack_from("<HHHHB", data, pos)
            pos += 9
            if f & 0x80:
                pos += 3 << ((f & 7) + 1)
            mcs, stream, pos = parse_gif_image_data(data, pos)
            frames.append((w, h, mcs, stream))
        else:
            raise ValueError(f"unexpected block 0x{tag:02x} in GIF data")
    return frames


# ---- TIFF -----------------------------------------------------------------------------------------
_TIFF_TYPES = {3: "H", 4: "I"}


def write_tiff_lzw(width: int, height: int, samples_per_pixel: int, rows_per_strip: int, strips,
                   predictor: int = 1) -> bytes:
    """A little-endian baseline TIFF (8 bits per sample, chunky) whose strips are the given
    TIFF-style LZW streams (Compression = 5).  predictor = 2 records that the strips were
    horizontally differenced before encoding (Codec.set_tiff_predictor / slzw_tiff_predictor_device
    with row_bytes = width * samples_per_pixel)."""
    if predictor not in (1, 2):
        raise ValueError("TIFF predictor must be 1 (none) or 2 (horizontal differencing)")
    n = len(strips)
    if n != (height + rows_per_strip - 1) // rows_per_strip:
        raise ValueError("number of strips does not match height / rows_per_strip")
    entries = []  # (tag, type, count, values)
    photometric = 2 if samples_per_pixel == 3 else 1
    entries.append((256, 4, 1, [width]))
    entries.append((257, 4, 1, [height]))
    entries.append((258, 3, samples_per_pixel, [8] * samples_per_pixel))
    entries.append((259, 3, 1, [5]))
    entries.append((262, 3, 1, [photometric]))
    entries.append((273, 4, n, None))  # StripOffsets, filled below
    entries.append((277, 3, 1, [samples_per_pixel]))
    entries.append((278, 4, 1, [rows_per_strip]))
    entries.append((279, 4, n, [len(s) for s in strips]))
    entries.append((284, 3, 1, [1]))
    if predictor == 2:
        entries.append((317, 3, 1, [2]))
    ifd_offset = 8
    ifd_size = 2 + 12 * len(entries) + 4
    extra_offset = ifd_offset + ifd_size
    extras = bytearray()
    placed = {}
    for tag, typ, count, values in entries:
        size = count * struct.calcsize(_TIFF_TYPES[typ])
        if size > 4:
            placed[tag] = extra_offset + len(extras)
            extras += b"\0" * size
            if len(extras) % 2:
                extras += b"\0"
    data_offset = extra_offset + len(extras)
    offsets, p = [], data_offset
    for s in strips:
        offsets.append(p)
        p += len(s) + (len(s) & 1)  # strips start on even offsets
    out = bytearray(b"II*\0" + struct.pack("<I", ifd_offset))
    out += struct.pack("<H", len(entries))
    for tag, typ, count, values in entries:
        if tag == 273:
            values = offsets
        fmt = "<" + _TIFF_TYPES[typ] * count
        raw = struct.pack(fmt, *values)
        if len(raw) <= 4:
            field = raw.ljust(4, b"\0")
        else:
            field = struct.pack("<I", placed[tag])
            extras[placed[tag] - extra_offset:placed[tag] - extra_offset + len(raw)] = raw
        out += struct.pack("<HHI", tag, typ, count) + field
    out += struct.pack("<I", 0)
    out += extras
    for s in strips:
        out += s
        if len(s) & 1:
            out += b"\0"
    return bytes(out)


def read_tiff(data: bytes) -> dict:
    """The first image of a TIFF file with Compression = 5 as a dict: width, height,
    samples_per_pixel, rows_per_strip, predictor (1 or 2), strips (list of bytes).  The strips go to
    TiffStyleDecoder / Codec.decode_batch; with predictor 2 the decoded strips still have to be
    accumulated (Codec.set_tiff_predictor(width * samples_per_pixel, samples_per_pixel))."""
    w, h, spp, rps, strips, predictor = _read_tiff(data)
    return {"width": w, "height": h, "samples_per_pixel": spp, "rows_per_strip": rps,
            "predictor": predictor, "strips": strips}


def read_tiff_strips(data: bytes):
    """(width, height, samples_per_pixel, rows_per_strip, [strip bytes]) of the first image of a
    TIFF file with Compression = 5 and no predictor."""
    w, h, spp, rps, strips, predictor = _read_tiff(data)
    if predictor != 1:
        raise ValueError("TIFF image uses a predictor: use read_tiff")
    return w, h, spp, rps, strips


def _read_tiff(data: bytes):
    if data[:2] == b"II":
        e = "<"
    elif data[:2] == b"MM":
        e = ">"
    else:
        raise ValueError("not a TIFF file")
    (ifd,) = struct.unpack_from(e + "I", data, 4)
    (count,) = struct.unpack_from(e + "H", data, ifd)
    tags = {}
    for i in range(count):
        tag, typ, cnt = struct.unpack_from(e + "HHI", data, ifd + 2 + 12 * i)
        if typ not in _TIFF_TYPES:
            continue
        fmt = e + _TIFF_TYPES[typ] * cnt
        size = struct.calcsize(fmt)
        off = ifd + 2 + 12 * i + 8
        if size > 4:
            (off,) = struct.unpack_from(e + "I", data, off)
        tags[tag] = list(struct.unpack_from(fmt, data, off))
    if tags.get(259, [1])[0] != 5:
        raise ValueError("TIFF image is not LZW-compressed")
    predictor = tags.get(317, [1])[0]
    if predictor not in (1, 2):
        raise ValueError("TIFF predictor %d is not supported" % predictor)
    if predictor == 2 and any(b != 8 for b in tags.get(258, [1])):
        raise ValueError("TIFF predictor 2 is supported for 8-bit samples only")
    offs, counts = tags[273], tags[279]
    strips = [data[o:o + c] for o, c in zip(offs, counts)]
    return (tags[256][0], tags[257][0], tags.get(277, [1])[0], tags.get(278, [tags[257][0]])[0], strips,
            predictor)


# ---- batches ---------------------------------------------------------------------------------------
def strips_of_image(pixels: np.ndarray, rows_per_strip: int):
    """(buffer, offsets) of a chunky 8-bit image cut into strips of rows_per_strip rows."""
    a = np.ascontiguousarray(pixels, dtype=np.uint8)
    h = a.shape[0]
    row = a.size // h
    bounds = [min(r, h) * row for r in range(0, h + rows_per_strip, rows_per_strip)]
    if bounds[-1] != h * row or (len(bounds) > 1 and bounds[-2] == bounds[-1]):
        bounds = sorted(set(bounds))
    return a.reshape(-1), np.asarray(bounds, dtype=np.uint64)


def split_dense(dense: np.ndarray, off: np.ndarray):
    """The streams of a dense batch as a list of bytes objects."""
    return [dense[int(off[i]):int(off[i + 1])].tobytes() for i in range(off.size - 1)]
