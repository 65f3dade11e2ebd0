"""Tiny encode/decode checks against the oracle (debug aid; every case is small)."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lzw_b200
from lzw_b200 import _lib
from oracle import oracle as O
from tests import cases as T

codec = lzw_b200.Codec(0)
rng = np.random.default_rng(1)
def gp(p): return _lib.Params(p.flavour, p.code_size, p.big_endian, p.tiff_early_change)
cases = [(O.gif(2), b"\x00\x00\x01\x03"), (O.tiff(), bytes(range(256)) * 2), (O.tiff(), T.make_stream(rng, "random", 3000, 255).tobytes()),
         (O.tiff(), T.make_stream(rng, "random", 20000, 255).tobytes()), (O.gif(2), T.make_stream(rng, "random", 30000, 3).tobytes()),
         (O.fixed(), T.make_stream(rng, "text", 40000, 255).tobytes()), (O.tiff(), T.make_stream(rng, "zeros", 100000, 255).tobytes())]
for i, (p, data) in enumerate(cases):
    print("case", i, T.pname(p), len(data), flush=True)
    got = codec.encode(gp(p), data)
    want = O.encode(p, data)
    print("   encode ok" if got == want else f"   ENCODE MISMATCH {got[0]} {len(got[2])} vs {want[0]} {len(want[2])}", flush=True)
print("done", flush=True)
