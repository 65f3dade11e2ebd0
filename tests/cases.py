"""Seeded inputs shared by the CPU and GPU test suites."""
import numpy as np

from oracle import oracle as O

D40 = bytes([1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 2, 2, 2, 2,
             2, 1, 1, 1, 0, 0, 0, 0, 2, 2, 2])
BAD_TIFF = bytes([0x1F, 0x40, 0x3A, 0x00, 0x00, 0x00, 0x44, 0x00, 0x00, 0x44, 0x00, 0x60, 0x54])


def all_variable_params():
    """Every (code_size, endianness, strategy) the Variable facade accepts."""
    return [O.variable(cs, be, tf) for cs in range(2, 9) for be in (False, True) for tf in (False, True)]


def all_params():
    return all_variable_params() + [O.fixed(False), O.fixed(True)]


def pname(p):
    if p.flavour == 1:
        return f"fixed-{'be' if p.big_endian else 'le'}"
    return f"var-cs{p.code_size}-{'be' if p.big_endian else 'le'}-{'tiff' if p.tiff_early_change else 'def'}"


def max_symbol(p):
    return 255 if p.flavour == 1 else (1 << p.code_size) - 1


def make_stream(rng, kind: str, n: int, hi: int) -> np.ndarray:
    """A stream of n symbols in [0, hi]."""
    if n == 0:
        return np.zeros(0, dtype=np.uint8)
    if kind == "random":
        return rng.integers(0, hi + 1, size=n, dtype=np.uint8)
    if kind == "zeros":
        return np.zeros(n, dtype=np.uint8)
    if kind == "runs":
        lens = rng.geometric(1 / 12, size=n // 4 + 4)
        vals = rng.integers(0, hi + 1, size=lens.size, dtype=np.uint8)
        return np.repeat(vals, lens)[:n].copy()
    if kind == "walk":
        steps = rng.integers(-2, 3, size=n)
        return (np.cumsum(steps) % (hi + 1)).astype(np.uint8)
    if kind == "text":
        p = 1.0 / np.arange(1, min(hi, 27) + 2) ** 1.1
        p /= p.sum()
        return rng.choice(p.size, size=n, p=p).astype(np.uint8)
    if kind == "period":
        base = rng.integers(0, hi + 1, size=int(rng.integers(1, 9)), dtype=np.uint8)
        return np.tile(base, n // base.size + 1)[:n].copy()
    raise ValueError(kind)


KINDS = ["random", "zeros", "runs", "walk", "text", "period"]


def make_batch(seed: int, n_streams: int, hi: int, max_len: int = 6000, with_empty: bool = True):
    """A ragged batch: (buf, off) with a mix of kinds and lengths, some empty / 1-byte streams."""
    rng = np.random.default_rng(seed)
    streams = []
    for i in range(n_streams):
        r = rng.random()
        if with_empty and r < 0.05:
            n = 0
        elif r < 0.12:
            n = int(rng.integers(1, 4))
        elif r < 0.6:
            n = int(rng.integers(4, 600))
        else:
            n = int(rng.integers(600, max_len))
        streams.append(make_stream(rng, KINDS[i % len(KINDS)], n, hi))
    off = np.zeros(n_streams + 1, dtype=np.uint64)
    off[1:] = np.cumsum([s.size for s in streams])
    buf = np.concatenate(streams) if streams else np.zeros(0, dtype=np.uint8)
    return buf.astype(np.uint8), off


def pack_dense(out, out_off, out_len):
    """Gathers the used part of every slot into a dense buffer + offsets."""
    n = out_len.size
    off = np.zeros(n + 1, dtype=np.uint64)
    off[1:] = np.cumsum(out_len)
    dense = np.empty(int(off[-1]), dtype=np.uint8)
    for i in range(n):
        l = int(out_len[i])
        dense[int(off[i]):int(off[i]) + l] = out[int(out_off[i]):int(out_off[i]) + l]
    return dense, off


def slots_equal(a, b, off, lens):
    """Compares the used part of every slot of two slotted buffers."""
    for i in range(lens.size):
        o, l = int(off[i]), int(lens[i])
        if not np.array_equal(a[o:o + l], b[o:o + l]):
            return i
    return -1
