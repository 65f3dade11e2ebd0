"""The lane-level model of the TIFF predictor kernels (tests/predictor_model.py) against the numpy
restatement of TIFF 6.0 section 14: unaligned strips, rows shorter than a lane's bytes, short last
rows, every samples-per-pixel value, and no byte outside the strip touched."""
import random

import numpy as np

from tests import predictor_model as M
from tests.test_predictor import hacc_ref, hdiff_ref


def test_predictor_kernel_model_matches_the_restatement():
    random.seed(7)
    rng = np.random.default_rng(7)
    for it in range(160):
        spp = random.choice([1, 2, 3, 4])
        pixels = random.randint(1, 400) if it % 3 else random.randint(1, 8)
        row_bytes = pixels * spp
        length = random.randint(0, 7000)
        addr = random.randint(16, 80)
        mem = rng.integers(0, 256, addr + length + 64, dtype=np.uint8)
        orig = mem.copy()
        want = orig.copy()
        want[addr:addr + length] = hdiff_ref(orig[addr:addr + length], row_bytes, spp)
        M.hdiff_strip(mem, addr, length, row_bytes, spp, threads=random.choice([2, 8, 32]))
        assert np.array_equal(mem, want), ("difference", it, spp, row_bytes, length, addr)
        M.hacc_strip(mem, addr, length, row_bytes, spp)
        assert np.array_equal(mem, orig), ("accumulate", it, spp, row_bytes, length, addr)
        mem = orig.copy()
        want = orig.copy()
        want[addr:addr + length] = hacc_ref(orig[addr:addr + length], row_bytes, spp)
        M.hacc_strip(mem, addr, length, row_bytes, spp)
        assert np.array_equal(mem, want), ("accumulate raw", it, spp, row_bytes, length, addr)
