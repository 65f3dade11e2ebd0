"""include/salzweg.hpp: the C++ mirror of the reference's public interface over the C ABI.
CPU: it compiles and links against libslzw.so.  GPU: the reference's own unit tests and doctests,
restated in tests/cpp/salzweg_kat.cpp, pass."""
import os
import subprocess

import pytest

from tests.conftest import GOLDEN, ROOT

CSRC = os.path.join(ROOT, "lzw_b200", "csrc")
EXE = os.path.join(ROOT, "tests", "cpp", "salzweg_kat")


EXE_COALESCE = os.path.join(ROOT, "tests", "cpp", "salzweg_coalesce")


def _build(name="salzweg_kat", exe=EXE):
    src = os.path.join(ROOT, "tests", "cpp", name + ".cpp")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-pthread", "-I", os.path.join(ROOT, "include"), src,
           "-o", exe, "-L", CSRC, "-lslzw", f"-Wl,-rpath,{CSRC}"]
    subprocess.check_call(cmd)


def test_cpp_facade_compiles_and_links():
    assert os.path.exists(os.path.join(CSRC, "libslzw.so")), "run __graft_entry__.build() first"
    _build()
    assert os.path.exists(EXE)
    _build("salzweg_coalesce", EXE_COALESCE)
    assert os.path.exists(EXE_COALESCE)


@pytest.mark.gpu
def test_cpp_facade_passes_the_reference_unit_tests():
    _build()
    r = subprocess.run([EXE, os.path.join(GOLDEN, "lorem_ipsum.txt"), os.path.join(GOLDEN, "lorem_ipsum_encoded.bin")],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout


@pytest.mark.gpu
def test_cpp_coalescing_facade_merges_concurrent_calls():
    """48 threads making one-stream calls through salzweg::coalesced get the plain
    functions' bytes and errors back while their calls are merged into batched launches."""
    _build("salzweg_coalesce", EXE_COALESCE)
    r = subprocess.run([EXE_COALESCE, os.path.join(GOLDEN, "lorem_ipsum.txt"),
                        os.path.join(GOLDEN, "lorem_ipsum_encoded.bin")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout
