"""The step algorithm of the fast decode kernel, modelled on the CPU (tests/fast_decode_model.py),
against the oracle: whenever the model finishes a stream itself, bytes, status and detail equal
the oracle's; everything else it defers.  Sized for the CPU suite (the full sweep is
`python tests/fast_decode_model.py`)."""
import numpy as np

from oracle import oracle as O
from tests import cases as T
from tests import fast_decode_model as M


def test_model_matches_oracle_on_clean_and_tight_streams():
    rng = np.random.default_rng(7)
    seen = {}
    for p in (O.tiff(), O.gif(2), O.gif(5), O.gif(8), O.fixed(False), O.fixed(True), O.variable(4, True, False)):
        hi = T.max_symbol(p)
        for kind in T.KINDS:
            for n in (0, 1, 5, 300, 6000):
                raw = T.make_stream(rng, kind, n, hi).tobytes()
                for cap in (None, n, n // 2):
                    r = M.check(p, raw, f"{T.pname(p)} {kind} {n} cap {cap}", cap=cap)
                    seen[r] = seen.get(r, 0) + 1
    assert seen.get("ok(status 0)", 0) > 300 and seen.get("ok(status 5)", 0) > 100, seen


def test_model_matches_oracle_on_corrupt_streams():
    rng = np.random.default_rng(8)
    seen = {}
    for p in (O.gif(8), O.tiff(), O.gif(3), O.fixed(False)):
        hi = T.max_symbol(p)
        for t in range(60):
            raw = T.make_stream(rng, T.KINDS[t % len(T.KINDS)], int(rng.integers(10, 2000)), hi).tobytes()

            def mut(pk):
                a = bytearray(pk)
                for _ in range(int(rng.integers(1, 4))):
                    if a:
                        a[int(rng.integers(0, len(a)))] ^= 1 << int(rng.integers(0, 8))
                if rng.random() < 0.3 and len(a) > 2:
                    del a[int(rng.integers(1, len(a))):]
                return bytes(a)
            r = M.check(p, raw, f"{T.pname(p)} corrupt {t}", mutate=mut, cap=len(raw) if t % 2 else None)
            seen[r] = seen.get(r, 0) + 1
    assert sum(v for k, v in seen.items() if k.startswith("ok")) > 150, seen


def test_model_long_words():
    for p in (O.tiff(), O.fixed(False)):
        raw = bytes(150_000)
        assert M.check(p, raw, "zeros").startswith("ok")
