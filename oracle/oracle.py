"""ctypes loader for the CPU oracle (oracle/slzw_oracle.c).

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; lzw_b200/ never does.

The oracle restates salzweg's encoder/decoder loops line by line
(/root/reference/lzw/src/{io,encoder,decoder}.rs); see the header of slzw_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

OK = 0
ERR_CODE_SIZE = 1
ERR_UNEXPECTED_CODE = 2
ERR_MISSING_CLEAR_CODE = 3
ERR_IO_UNEXPECTED_EOF = 4
ERR_IO_WRITE_ZERO = 5
ERR_REFERENCE_PANIC = 6


class Params(C.Structure):
    _fields_ = [
        ("flavour", C.c_uint8),
        ("code_size", C.c_uint8),
        ("big_endian", C.c_uint8),
        ("tiff_early_change", C.c_uint8),
    ]


class Batch(C.Structure):
    _fields_ = [
        ("in_", C.c_void_p),
        ("in_off", C.c_void_p),
        ("out", C.c_void_p),
        ("out_off", C.c_void_p),
        ("out_len", C.c_void_p),
        ("status", C.c_void_p),
        ("detail", C.c_void_p),
        ("code_size", C.c_void_p),
        ("n", C.c_uint64),
    ]


def gif(code_size: int) -> Params:
    """GifStyleEncoder/Decoder preset (encoder.rs:392-399, decoder.rs:333-340)."""
    return Params(0, code_size, 0, 0)


def tiff() -> Params:
    """TiffStyleEncoder/Decoder preset (encoder.rs:479-487, decoder.rs:420-428)."""
    return Params(0, 8, 1, 1)


def fixed(big_endian: bool = False) -> Params:
    """FixedEncoder/Decoder preset (encoder.rs:565-576, decoder.rs:503-514)."""
    return Params(1, 0, 1 if big_endian else 0, 0)


def variable(code_size: int, big_endian: bool, tiff_early_change: bool) -> Params:
    """VariableEncoder/Decoder (encoder.rs:199-220, decoder.rs:99-120)."""
    return Params(0, code_size, 1 if big_endian else 0, 1 if tiff_early_change else 0)


def lenient(code_size: int, big_endian: bool = False, tiff_early_change: bool = False) -> Params:
    """Extension of include/slzw.h (not in the reference): a full dictionary freezes, no MissingClearCode."""
    return Params(2, code_size, 1 if big_endian else 0, 1 if tiff_early_change else 0)


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (building the checker is not using it)."""
    src = os.path.join(_HERE, "slzw_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        for name in ("oracle_encode", "oracle_decode"):
            f = getattr(_lib, name)
            f.restype = C.c_int
            f.argtypes = [C.POINTER(Params), C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                          C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
        for name in ("oracle_encode_batch", "oracle_decode_batch"):
            f = getattr(_lib, name)
            f.restype = C.c_int
            f.argtypes = [C.POINTER(Params), C.POINTER(Batch), C.c_int]
        _lib.oracle_bitwrite.restype = C.c_uint64
        _lib.oracle_bitwrite.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        _lib.oracle_bitread.restype = C.c_uint64
        _lib.oracle_bitread.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p]
    return _lib


def encode_bound(n: int) -> int:
    """A safe output size for an n-byte stream (every byte a 12-bit code + clears + EOI)."""
    return (n + 2 + n // 3000 + 2) * 3 // 2 + 8


def _as_u8(data) -> np.ndarray:
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data, dtype=np.uint8)
    return np.frombuffer(bytes(data), dtype=np.uint8)


def encode(params: Params, data, cap: int | None = None):
    """Returns (status, detail, bytes_produced)."""
    a = _as_u8(data)
    cap = encode_bound(a.size) if cap is None else cap
    out = np.empty(max(cap, 1), dtype=np.uint8)
    out_len = C.c_uint64(0)
    detail = C.c_uint32(0)
    st = lib().oracle_encode(C.byref(params), a.ctypes.data, a.size, out.ctypes.data, cap,
                             C.byref(out_len), C.byref(detail))
    return st, detail.value, out[: out_len.value].tobytes()


def decoded_size(params: Params, data):
    """Size-only decode: (status, detail, bytes the reference would have produced)."""
    a = _as_u8(data)
    out_len = C.c_uint64(0)
    detail = C.c_uint32(0)
    st = lib().oracle_decode(C.byref(params), a.ctypes.data, a.size, None, 0,
                             C.byref(out_len), C.byref(detail))
    return st, detail.value, out_len.value


def decode(params: Params, data, cap: int | None = None):
    """Returns (status, detail, bytes_produced).  cap=None sizes the output first (Vec)."""
    a = _as_u8(data)
    if cap is None:
        _, _, cap = decoded_size(params, a)
    out = np.empty(max(cap, 1), dtype=np.uint8)
    out_len = C.c_uint64(0)
    detail = C.c_uint32(0)
    st = lib().oracle_decode(C.byref(params), a.ctypes.data, a.size, out.ctypes.data, cap,
                             C.byref(out_len), C.byref(detail))
    return st, detail.value, out[: out_len.value].tobytes()


def _batch(params: Params, decode_: bool, in_buf, in_off, out_off, code_size, threads: int,
           write_out: bool = True):
    in_buf = _as_u8(in_buf)
    in_off = np.ascontiguousarray(in_off, dtype=np.uint64)
    n = in_off.size - 1
    out_off = np.ascontiguousarray(out_off, dtype=np.uint64)
    out = np.empty(max(int(out_off[-1]), 1), dtype=np.uint8) if write_out else None
    out_len = np.zeros(n, dtype=np.uint64)
    status = np.zeros(n, dtype=np.uint32)
    detail = np.zeros(n, dtype=np.uint32)
    cs = None if code_size is None else np.ascontiguousarray(code_size, dtype=np.uint8)
    b = Batch(in_buf.ctypes.data, in_off.ctypes.data, out.ctypes.data if write_out else None,
              out_off.ctypes.data, out_len.ctypes.data, status.ctypes.data, detail.ctypes.data,
              None if cs is None else cs.ctypes.data, n)
    f = lib().oracle_decode_batch if decode_ else lib().oracle_encode_batch
    f(C.byref(params), C.byref(b), threads)
    return out, out_len, status, detail


def encode_batch(params, in_buf, in_off, out_off, code_size=None, threads=1):
    """Batch encode with the slzw_batch layout; returns (out, out_len, status, detail)."""
    return _batch(params, False, in_buf, in_off, out_off, code_size, threads)


def decode_batch(params, in_buf, in_off, out_off, code_size=None, threads=1, write_out=True):
    return _batch(params, True, in_buf, in_off, out_off, code_size, threads, write_out)


def bitwrite(big: bool, codes, widths) -> bytes:
    codes = np.ascontiguousarray(codes, dtype=np.uint16)
    widths = np.ascontiguousarray(widths, dtype=np.uint8)
    out = np.empty(codes.size * 2 + 8, dtype=np.uint8)
    n = lib().oracle_bitwrite(int(big), codes.ctypes.data, widths.ctypes.data, codes.size,
                              out.ctypes.data, out.size)
    return out[:n].tobytes()


def bitread(big: bool, data, widths):
    a = _as_u8(data)
    widths = np.ascontiguousarray(widths, dtype=np.uint8)
    codes = np.zeros(widths.size, dtype=np.uint16)
    n = lib().oracle_bitread(int(big), a.ctypes.data, a.size, widths.ctypes.data, widths.size,
                             codes.ctypes.data)
    return [int(c) for c in codes[:n]]
