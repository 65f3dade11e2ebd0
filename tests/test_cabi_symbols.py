"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol the header
declares, and fails loudly (no fallback) when there is no GPU.  No compute calls here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    with open(os.path.join(ROOT, "include", "slzw.h")) as f:
        text = f.read()
    return sorted(set(re.findall(r"SLZW_API[^;(]*?\b(slzw_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from lzw_b200 import _lib
    declared = header_symbols()
    assert len(declared) >= 17
    lib = _lib.lib()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/slzw.h but not exported"
    # the ctypes table binds exactly the declared set
    assert sorted(_lib.SYMBOLS) == declared


def test_status_messages_match_reference_display_strings():
    from lzw_b200.codec import status_message
    # encoder.rs:35-41 and decoder.rs:31-39; the encoder's CodeSize text ends with a period
    assert status_message(False, 1, 10) == "Code size must be between 2 and 8, was 10."
    assert status_message(True, 1, 10) == "Code size must be between 2 and 8, was 10"
    assert status_message(False, 2, 8, 2) == "Unexpected code 8. For code size 2, data should be < 4."
    assert status_message(True, 2, 258) == "Unexpected code while decompressing: 258"
    assert status_message(True, 3, 0) == "Dictionnary growing past 4096, expected CLEAR_CODE missing"
    assert status_message(True, 4, 0) == "failed to fill whole buffer"
    assert status_message(False, 5, 0) == "failed to write whole buffer"


def test_encode_bound_covers_oracle_worst_cases():
    import numpy as np
    from lzw_b200 import _lib
    from oracle import oracle as O
    lib = _lib.lib()
    rng = np.random.default_rng(0)
    for p in (O.tiff(), O.gif(8), O.gif(2), O.fixed(False)):
        gp = _lib.Params(p.flavour, p.code_size, p.big_endian, p.tiff_early_change)
        hi = 255 if p.flavour else (1 << p.code_size) - 1
        for n in (0, 1, 2, 100, 5000, 70000):
            data = rng.integers(0, hi + 1, size=n, dtype=np.uint8)
            st, _, out = O.encode(p, data, cap=n * 2 + 64)
            assert st == 0 and len(out) <= lib.slzw_encode_bound(C.byref(gp), n)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import lzw_b200
    from lzw_b200.codec import SlzwError
    with pytest.raises(SlzwError, match="no CPU fallback"):
        lzw_b200.Codec(0)
    from lzw_b200.encoder import GifStyleEncoder
    with pytest.raises(SlzwError):
        GifStyleEncoder.encode_to_vec(b"\x00\x01", 2)


def test_product_never_imports_oracle():
    """The product path must not route through oracle/ (tier rule 3)."""
    pkg = os.path.join(ROOT, "lzw_b200")
    for dirpath, _, files in os.walk(pkg):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                with open(os.path.join(dirpath, name)) as f:
                    text = f.read()
                assert "oracle" not in text.lower(), f"{name} mentions the oracle"
