"""Synthetic workloads of BASELINE.json's configs, as (buffer, offsets) batches.

Every stream draws from its own SplitMix64 generator seeded from (seed, stream index) alone
(tools/wlgen/wlgen.c: integer-only, reproducible from C / C++ / Rust), so a batch of n streams is a
prefix of any longer batch and rank r of a sharded run can generate exactly its own streams.
Streams are laid out back to back in one uint8 buffer; `off` has n+1 uint64 entries.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_GOLDEN = os.path.join(os.path.dirname(_HERE), "tests", "golden")

SEED = 0x5A172E60

_WL = None


def _wl():
    """tools/wlgen/libwlgen.so (built by __graft_entry__.build()): the SplitMix64 generators."""
    global _WL
    if _WL is None:
        path = os.path.join(os.path.dirname(_HERE), "tools", "wlgen", "libwlgen.so")
        if not os.path.exists(path):  # a fresh checkout that skipped build(): one gcc call
            import subprocess
            subprocess.check_call(["make", "-C", os.path.dirname(path)])
        L = ctypes.CDLL(path)
        u64, vp = ctypes.c_uint64, ctypes.c_void_p
        L.wl_tiff_strip_lens.argtypes = [u64, u64, u64, u64, u64, vp]
        L.wl_tiff_strips_fill.argtypes = [u64, u64, u64, u64, u64, vp, vp]
        L.wl_gif_frames_fill.argtypes = [u64, u64, u64, u64, vp]
        L.wl_text_chunks_fill.argtypes = [u64, u64, u64, u64, vp, vp, ctypes.c_uint32, vp]
        for f in (L.wl_tiff_strip_lens, L.wl_tiff_strips_fill, L.wl_gif_frames_fill, L.wl_text_chunks_fill):
            f.restype = None
        _WL = L
    return _WL


def tiff_strips(n_streams: int = 65536, seed: int = SEED + 3, lo: int = 8192, span: int = 57345,
                first: int = 0):
    """Config 3: TIFF strips of 8-64 KB, entropy class i mod 4:
    0 uniform random bytes, 1 photo-like random walk, 2 runs (mean 16), 3 16-symbol Zipf(1.2).
    Streams [first, first + n_streams) of the (infinite) sequence that `seed` names."""
    L = _wl()
    lens = np.empty(n_streams, dtype=np.uint64)
    L.wl_tiff_strip_lens(seed, first, n_streams, lo, span, lens.ctypes.data)
    off = np.zeros(n_streams + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    buf = np.empty(int(off[-1]), dtype=np.uint8)
    L.wl_tiff_strips_fill(seed, first, n_streams, lo, span, off.ctypes.data, buf.ctypes.data)
    return buf, off


def gif_frames(n_frames: int = 4096, side: int = 1024, seed: int = SEED + 4, first: int = 0):
    """Config 4: 8-bit palette frames, cs_i = 2 + (i mod 7), pixels < 2^cs; runs (mean 8) with about
    10 % per-pixel dither noise; every 8th frame is pure noise.  Returns (buf, off, code_size)."""
    px = side * side
    off = (np.arange(n_frames + 1, dtype=np.uint64) * np.uint64(px))
    buf = np.empty(n_frames * px, dtype=np.uint8)
    code_size = (2 + ((first + np.arange(n_frames)) % 7)).astype(np.uint8)
    _wl().wl_gif_frames_fill(seed, first, n_frames, px, buf.ctypes.data)
    return buf, off, code_size


_WORDS = None


def _lorem_words():
    """Sorted unique lower-case words of tests/golden/lorem_ipsum.txt as (blob, offsets)."""
    global _WORDS
    if _WORDS is None:
        with open(os.path.join(_GOLDEN, "lorem_ipsum.txt"), "rb") as f:
            raw = f.read()
        words = sorted({w.strip(b".,;:!?").lower() for w in raw.split() if w.strip(b".,;:!?")})
        woff = np.zeros(len(words) + 1, dtype=np.uint32)
        woff[1:] = np.cumsum([len(w) for w in words])
        _WORDS = (np.frombuffer(b"".join(words), dtype=np.uint8).copy(), woff)
    return _WORDS


def text_chunks(n_chunks: int = 131072, chunk: int = 65536, seed: int = SEED + 5, first: int = 0,
                corpus_bytes: int = 0):
    """Config 5: lorem-like text (words of lorem_ipsum.txt drawn uniformly, sentences of 12+ words,
    80-column lines) in fixed-size chunks, every chunk from its own generator."""
    blob, woff = _lorem_words()
    buf = np.empty(n_chunks * chunk, dtype=np.uint8)
    _wl().wl_text_chunks_fill(seed, first, n_chunks, chunk, blob.ctypes.data, woff.ctypes.data,
                              woff.size - 1, buf.ctypes.data)
    off = np.arange(n_chunks + 1, dtype=np.uint64) * np.uint64(chunk)
    return buf, off


def sunflower_strips(strip: int = 8192):
    """Config 2: pixel bytes of sunflower.bmp (54-byte header skipped) cut into 8 KB strips."""
    with open(os.path.join(_GOLDEN, "sunflower.bmp"), "rb") as f:
        px = np.frombuffer(f.read()[54:], dtype=np.uint8).copy()
    bounds = list(range(0, px.size, strip)) + [px.size]
    return px, np.asarray(bounds, dtype=np.uint64)


def encode_slots(off: np.ndarray, align: int = 16) -> np.ndarray:
    """Worst-case output slots (slzw_encode_bound per stream, rounded up to `align`)."""
    lens = np.diff(np.asarray(off, dtype=np.uint64)).astype(np.uint64)
    slots = ((lens + np.uint64(3) + lens // np.uint64(3838) + np.uint64(1)) * np.uint64(12) + np.uint64(7)) // np.uint64(8)
    slots = (slots + np.uint64(align - 1)) // np.uint64(align) * np.uint64(align)
    out = np.zeros(lens.size + 1, dtype=np.uint64)
    out[1:] = np.cumsum(slots)
    return out
