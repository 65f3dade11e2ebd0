"""Synthetic workloads of BASELINE.json's configs, as (buffer, offsets) batches.

Every generator is deterministic in (seed, n_streams).  Streams are laid out back to back in
one uint8 buffer; `off` has n+1 uint64 entries.  The generators are numpy-only (host); bench.py
uploads the result once, outside the timed region.
"""
from __future__ import annotations

import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_GOLDEN = os.path.join(os.path.dirname(_HERE), "tests", "golden")

SEED = 0x5A172E60


def _lengths(rng, n, lo, span):
    return (lo + rng.integers(0, span, size=n, dtype=np.int64)).astype(np.int64)


def _pool_random(rng, total):
    return rng.integers(0, 256, size=total, dtype=np.uint8)


def _pool_photo(rng, total):
    """3-channel interleaved random walk, step in [-3, 3], wrapping mod 256."""
    px = (total + 2) // 3
    steps = rng.integers(-3, 4, size=(px, 3), dtype=np.int8).view(np.uint8)
    start = rng.integers(0, 256, size=(1, 3), dtype=np.uint8)
    walk = np.cumsum(steps, axis=0, dtype=np.uint8) + start
    return walk.reshape(-1)[:total]


def _pool_runs(rng, total, mean_run=16, alphabet=256):
    nruns = int(total / mean_run * 1.1) + 16
    out = np.empty(0, dtype=np.uint8)
    while out.size < total:
        lens = rng.geometric(1.0 / mean_run, size=nruns)
        vals = rng.integers(0, alphabet, size=nruns, dtype=np.uint8)
        out = np.concatenate([out, np.repeat(vals, lens)])
    return out[:total]


def _pool_zipf(rng, total, symbols=16, s=1.2):
    p = 1.0 / np.arange(1, symbols + 1) ** s
    p /= p.sum()
    cdf = np.cumsum(p)
    alphabet = np.frombuffer(b" etaoinshrdlucmf", dtype=np.uint8)[:symbols]
    u = rng.random(total, dtype=np.float32)
    return alphabet[np.minimum(np.searchsorted(cdf, u), symbols - 1)]


def tiff_strips(n_streams: int = 65536, seed: int = SEED + 3, lo: int = 8192, span: int = 57345):
    """Config 3: TIFF strips of 8-64 KB, entropy class i mod 4:
    0 uniform random bytes, 1 photo-like random walk, 2 runs (mean 16), 3 16-symbol Zipf(1.2)."""
    rng = np.random.default_rng(seed)
    lens = _lengths(rng, n_streams, lo, span)
    off = np.zeros(n_streams + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    buf = np.empty(int(off[-1]), dtype=np.uint8)
    makers = [_pool_random, _pool_photo, _pool_runs, _pool_zipf]
    for cls in range(4):
        idx = np.arange(cls, n_streams, 4)
        if idx.size == 0:
            continue
        total = int(lens[idx].sum())
        pool = makers[cls](rng, total)
        p = 0
        for i in idx:
            l = int(lens[i])
            o = int(off[i])
            buf[o:o + l] = pool[p:p + l]
            p += l
    return buf, off


def gif_frames(n_frames: int = 4096, side: int = 1024, seed: int = SEED + 4):
    """Config 4: 8-bit palette frames, cs_i = 2 + (i mod 7), pixels < 2^cs; rows made of runs
    (mean 8) with 10 % per-pixel dither noise; every 8th frame is pure noise.
    Returns (buf, off, code_size)."""
    rng = np.random.default_rng(seed)
    px = side * side
    off = (np.arange(n_frames + 1, dtype=np.uint64) * np.uint64(px))
    buf = np.empty(n_frames * px, dtype=np.uint8)
    code_size = (2 + (np.arange(n_frames) % 7)).astype(np.uint8)
    for i in range(n_frames):
        colours = 1 << int(code_size[i])
        if i % 8 == 7:
            frame = rng.integers(0, colours, size=px, dtype=np.uint8)
        else:
            frame = _pool_runs(rng, px, mean_run=8, alphabet=colours)
            noisy = rng.random(px, dtype=np.float32) < 0.10
            frame = np.where(noisy, rng.integers(0, colours, size=px, dtype=np.uint8), frame)
        buf[i * px:(i + 1) * px] = frame
    return buf, off, code_size


def lorem_corpus(nbytes: int, seed: int = SEED + 5) -> np.ndarray:
    """Lorem-like text: words drawn uniformly from the word list of lorem_ipsum.txt, separated
    by spaces, '. ' + capital about every 12 words, newline about every 80 characters."""
    with open(os.path.join(_GOLDEN, "lorem_ipsum.txt"), "rb") as f:
        raw = f.read()
    words = sorted({w.strip(b".,;:!?").lower() for w in raw.split() if w.strip(b".,;:!?")})
    rng = np.random.default_rng(seed)
    out = bytearray()
    line = 0
    since_stop = 0
    cap = True
    while len(out) < nbytes:
        for wi in rng.integers(0, len(words), size=4096):
            w = words[int(wi)]
            if cap:
                w = w[:1].upper() + w[1:]
                cap = False
            out += w
            line += len(w)
            since_stop += 1
            if since_stop >= 12 and rng.random() < 0.5:
                out += b"."
                since_stop = 0
                cap = True
            if line >= 80:
                out += b"\n"
                line = 0
            else:
                out += b" "
                line += 1
    return np.frombuffer(bytes(out[:nbytes]), dtype=np.uint8).copy()


def text_chunks(n_chunks: int = 131072, chunk: int = 65536, seed: int = SEED + 5,
                corpus_bytes: int = 32 << 20):
    """Config 5: lorem-like text in fixed-size chunks.  Each chunk is a window of a
    `corpus_bytes` corpus at a pseudo-random offset (chunks are encoded independently, so
    overlap between windows does not couple them)."""
    corpus = lorem_corpus(max(corpus_bytes, 2 * chunk), seed)
    rng = np.random.default_rng(seed + 1)
    starts = rng.integers(0, corpus.size - chunk, size=n_chunks)
    buf = np.empty(n_chunks * chunk, dtype=np.uint8)
    for i, s in enumerate(starts):
        buf[i * chunk:(i + 1) * chunk] = corpus[s:s + chunk]
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
