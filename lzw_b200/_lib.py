"""ctypes binding of libslzw.so (the C ABI declared in include/slzw.h).

There is no fallback: if the CUDA library is missing or no sm_100 device is usable, importing
works (so CPU-only tooling can inspect symbols) but creating a context raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SLZW_LIB: another build of the same library (tuning experiments, tools/build_variants.sh)
LIB_PATH = os.environ.get("SLZW_LIB") or os.path.join(_HERE, "csrc", "libslzw.so")

# slzw_status
OK = 0
ERR_CODE_SIZE = 1
ERR_UNEXPECTED_CODE = 2
ERR_MISSING_CLEAR_CODE = 3
ERR_IO_UNEXPECTED_EOF = 4
ERR_IO_WRITE_ZERO = 5
ERR_REFERENCE_PANIC = 6

# slzw_rc
RC_OK = 0
RC_CUDA = -1
RC_INVALID = -2
RC_NO_DEVICE = -3
RC_NOMEM = -4

FLAVOUR_VARIABLE = 0
FLAVOUR_FIXED = 1
FLAVOUR_VARIABLE_LENIENT = 2


class Params(C.Structure):
    """slzw_params"""
    _fields_ = [
        ("flavour", C.c_uint8),
        ("code_size", C.c_uint8),
        ("big_endian", C.c_uint8),
        ("tiff_early_change", C.c_uint8),
    ]

    def __repr__(self):
        return (f"Params(flavour={self.flavour}, code_size={self.code_size}, "
                f"big_endian={self.big_endian}, tiff_early_change={self.tiff_early_change})")


class Batch(C.Structure):
    """slzw_batch"""
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


# every symbol include/slzw.h declares: name -> (restype, argtypes)
_P = C.POINTER
SYMBOLS = {
    "slzw_create": (C.c_int, [C.c_int, _P(C.c_void_p)]),
    "slzw_destroy": (None, [C.c_void_p]),
    "slzw_last_error": (C.c_char_p, [C.c_void_p]),
    "slzw_kernel_launches": (C.c_uint64, [C.c_void_p]),
    "slzw_version": (C.c_uint32, []),
    "slzw_last_deferred": (C.c_uint64, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "slzw_last_encode_shares": (C.c_int, [C.c_void_p, C.c_void_p]),
    "slzw_encode_batch_device": (C.c_int, [C.c_void_p, _P(Params), _P(Batch), C.c_void_p]),
    "slzw_decode_batch_device": (C.c_int, [C.c_void_p, _P(Params), _P(Batch), C.c_void_p]),
    "slzw_encode_batch_host": (C.c_int, [C.c_void_p, _P(Params), _P(Batch)]),
    "slzw_decode_batch_host": (C.c_int, [C.c_void_p, _P(Params), _P(Batch)]),
    "slzw_encode_batch_host_dense": (C.c_int, [C.c_void_p, _P(Params), C.c_void_p, C.c_void_p,
                                               C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p,
                                               C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                               _P(C.c_uint64)]),
    "slzw_encode_batch_host_dense_begin": (C.c_int, [C.c_void_p, _P(Params), C.c_void_p, C.c_void_p,
                                                     C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p,
                                                     C.c_void_p, C.c_void_p, _P(C.c_uint64)]),
    "slzw_encode_batch_host_dense_finish": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "slzw_multi_create": (C.c_int, [C.c_void_p, C.c_int, _P(C.c_void_p)]),
    "slzw_multi_destroy": (None, [C.c_void_p]),
    "slzw_multi_device_count": (C.c_int, [C.c_void_p]),
    "slzw_multi_last_error": (C.c_char_p, [C.c_void_p]),
    "slzw_multi_kernel_launches": (C.c_uint64, [C.c_void_p]),
    "slzw_partition_streams": (None, [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]),
    "slzw_multi_encode_batch_host": (C.c_int, [C.c_void_p, _P(Params), _P(Batch)]),
    "slzw_multi_decode_batch_host": (C.c_int, [C.c_void_p, _P(Params), _P(Batch)]),
    "slzw_multi_encode_batch_host_dense": (C.c_int, [C.c_void_p, _P(Params), C.c_void_p, C.c_void_p,
                                                     C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p,
                                                     C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                                     _P(C.c_uint64)]),
    "slzw_encode": (C.c_int, [C.c_void_p, _P(Params), C.c_void_p, C.c_uint64, C.c_void_p,
                              C.c_uint64, _P(C.c_uint64), _P(C.c_uint32)]),
    "slzw_decode": (C.c_int, [C.c_void_p, _P(Params), C.c_void_p, C.c_uint64, C.c_void_p,
                              C.c_uint64, _P(C.c_uint64), _P(C.c_uint32)]),
    "slzw_encode_bound": (C.c_uint64, [_P(Params), C.c_uint64]),
    "slzw_decoded_sizes_batch_device": (C.c_int, [C.c_void_p, _P(Params), _P(Batch), C.c_void_p]),
    "slzw_compact_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                      C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "slzw_decoded_sizes_batch_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "slzw_tiff_predictor_device": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p]),
    "slzw_set_tiff_predictor": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32]),
    "slzw_host_alloc": (C.c_void_p, [C.c_size_t]),
    "slzw_host_free": (None, [C.c_void_p]),
    "slzw_status_message": (C.c_int, [C.c_int, C.c_uint32, C.c_uint32, C.c_uint8, C.c_char_p,
                                      C.c_size_t]),
}

_lib = None


def lib() -> C.CDLL:
    """Loads libslzw.so; raises if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA library has not been built "
                "(run `make -C lzw_b200/csrc` or `python -c 'import __graft_entry__ as g; g.build()'`); "
                "lzw_b200 has no CPU fallback")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(l, name)  # AttributeError if the header and the library disagree
            f.restype = res
            f.argtypes = args
        _lib = l
    return _lib
