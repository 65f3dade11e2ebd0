"""include/salzweg.hpp: the C++ mirror of the reference's public interface over the C ABI.
CPU: it compiles and links against libslzw.so.  GPU: the reference's own unit tests and doctests,
restated in tests/cpp/salzweg_kat.cpp, pass."""
import os
import subprocess

import pytest

from tests.conftest import GOLDEN, ROOT

CSRC = os.path.join(ROOT, "lzw_b200", "csrc")
EXE = os.path.join(ROOT, "tests", "cpp", "salzweg_kat")


def _build():
    src = os.path.join(ROOT, "tests", "cpp", "salzweg_kat.cpp")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
           "-L", CSRC, "-lslzw", f"-Wl,-rpath,{CSRC}"]
    subprocess.check_call(cmd)


def test_cpp_facade_compiles_and_links():
    assert os.path.exists(os.path.join(CSRC, "libslzw.so")), "run __graft_entry__.build() first"
    _build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_facade_passes_the_reference_unit_tests():
    _build()
    r = subprocess.run([EXE, os.path.join(GOLDEN, "lorem_ipsum.txt"), os.path.join(GOLDEN, "lorem_ipsum_encoded.bin")],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout
