"""Randomised parity sweep on the GPU: every flavour, per-stream code sizes, ragged lengths, tight and
generous output slots, clean and corrupted inputs, both through the host entry points (chunked
pipeline with a random chunk size) -- sizes, statuses, details and bytes equal the oracle's."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests import cases as T

pytestmark = pytest.mark.gpu


def gp(p):
    from lzw_b200 import _lib
    return _lib.Params(p.flavour, p.code_size, p.big_endian, p.tiff_early_change)


def _corrupt(rng, packed: np.ndarray) -> np.ndarray:
    a = packed.copy()
    for _ in range(int(rng.integers(1, 4))):
        if a.size:
            a[int(rng.integers(0, a.size))] ^= np.uint8(1 << int(rng.integers(0, 8)))
    if rng.random() < 0.3 and a.size > 2:
        a = a[: int(rng.integers(1, a.size))]
    return a


@pytest.mark.parametrize("seed", range(int(os.environ.get("SLZW_FUZZ_SEEDS", "6"))))  # more seeds: export SLZW_FUZZ_SEEDS
def test_random_batches_match_the_oracle(seed, monkeypatch):
    import lzw_b200
    rng = np.random.default_rng(1000 + seed)
    monkeypatch.setenv("SLZW_HOST_CHUNK_BYTES", str(int(rng.integers(2_000, 400_000))))
    # odd seeds force the latency variant of the encoder where it exists (fixed flavour)
    monkeypatch.setenv("SLZW_ENC_CONFIG", "1" if seed % 2 else "0")
    codec = lzw_b200.Codec(0)
    try:
        for p in (O.tiff(), O.gif(int(rng.integers(2, 9))), O.fixed(bool(rng.integers(0, 2))),
                  O.variable(int(rng.integers(2, 9)), bool(rng.integers(0, 2)), bool(rng.integers(0, 2)))):
            n = int(rng.integers(20, 260))
            per_stream_cs = p.flavour == 0 and rng.random() < 0.5
            cs = rng.integers(2, 9, size=n).astype(np.uint8) if per_stream_cs else None
            streams = []
            for i in range(n):
                hi = (1 << int(cs[i])) - 1 if cs is not None else T.max_symbol(p)
                if rng.random() < 0.05:
                    hi = 255  # out-of-range symbols: UnexpectedCode / first-byte cases
                length = int(rng.choice([0, 1, 2, 3, 17, 300, 3000, 9000, 40000], p=[.04, .04, .04, .04, .1, .3, .3, .1, .04]))
                streams.append(T.make_stream(rng, T.KINDS[int(rng.integers(0, len(T.KINDS)))], length, hi))
            off = np.zeros(n + 1, dtype=np.uint64)
            off[1:] = np.cumsum([s.size for s in streams])
            buf = np.concatenate(streams) if off[-1] else np.zeros(0, dtype=np.uint8)
            # output slots: mostly generous, some too small
            slot = np.array([O.encode_bound(int(s.size)) if rng.random() < 0.9 else int(rng.integers(0, 40))
                             for s in streams], dtype=np.uint64)
            slots = np.zeros(n + 1, dtype=np.uint64)
            slots[1:] = np.cumsum(slot)
            out, _, out_len, st, det = codec.encode_batch(gp(p), buf, off, out_off=slots, code_size=cs)
            o_out, o_len, o_st, o_det = O.encode_batch(p, buf, off, slots, code_size=cs)
            assert np.array_equal(st, o_st), np.nonzero(st != o_st)[0][:5]
            assert np.array_equal(out_len, o_len) and np.array_equal(det, o_det)
            assert T.slots_equal(out, o_out, slots, o_len) == -1
            # decode what the oracle produced, partly corrupted, into exact / tight / generous slots
            parts = []
            for i in range(n):
                e = o_out[int(slots[i]):int(slots[i]) + int(o_len[i])]
                parts.append(_corrupt(rng, e) if rng.random() < 0.25 else e)
            d_off = np.zeros(n + 1, dtype=np.uint64)
            d_off[1:] = np.cumsum([x.size for x in parts])
            d_in = np.concatenate(parts) if d_off[-1] else np.zeros(0, dtype=np.uint8)
            room = np.array([s.size if rng.random() < 0.6 else int(s.size * rng.random() * 2) for s in streams],
                            dtype=np.uint64)
            caps = np.zeros(n + 1, dtype=np.uint64)
            caps[1:] = np.cumsum(room)
            dec, dlen, dst, ddet = codec.decode_batch(gp(p), d_in, d_off, caps, code_size=cs)
            o_dec, o_dlen, o_dst, o_ddet = O.decode_batch(p, d_in, d_off, caps, code_size=cs)
            assert np.array_equal(dst, o_dst), (T.pname(p), np.nonzero(dst != o_dst)[0][:5], dst[dst != o_dst][:5], o_dst[dst != o_dst][:5])
            assert np.array_equal(dlen, o_dlen) and np.array_equal(ddet, o_ddet)
            assert T.slots_equal(dec, o_dec, caps, o_dlen) == -1
    finally:
        codec.close()
