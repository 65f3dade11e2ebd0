"""Host-side mirror of lzw/src/encoder.rs: the four encoder types and EncodingError.

Same names, argument meaning and error behaviour as the reference; the work is done on the
GPU by libslzw (there is no CPU path).  `data` is anything convertible to bytes or an object
with .read(); `into` is anything with .write() (a file, io.BytesIO) or a bytearray.
"""
from __future__ import annotations

from . import _lib
from .codec import default_codec, status_message
from .types import (CodeSizeStrategy, Endianness, fixed_params, gif_params, tiff_params,
                    variable_params)


class EncodingError(Exception):
    """encoder.rs:16-29.  `kind` is 'Io', 'CodeSize' or 'UnexpectedCode'."""

    def __init__(self, kind: str, message: str, code: int | None = None,
                 code_size: int | None = None, io_kind: str | None = None):
        super().__init__(message)
        self.kind, self.code, self.code_size, self.io_kind = kind, code, code_size, io_kind

    @staticmethod
    def CodeSize(code_size: int) -> "EncodingError":
        return EncodingError("CodeSize", status_message(False, _lib.ERR_CODE_SIZE, code_size),
                             code_size=code_size)

    @staticmethod
    def UnexpectedCode(code: int, code_size: int) -> "EncodingError":
        return EncodingError("UnexpectedCode",
                             status_message(False, _lib.ERR_UNEXPECTED_CODE, code, code_size),
                             code=code, code_size=code_size)


def _read_all(data) -> bytes:
    return data.read() if hasattr(data, "read") else bytes(data)


def _write(into, payload: bytes):
    if hasattr(into, "write"):
        into.write(payload)
    else:
        into.extend(payload)


def _raise(status: int, detail: int, code_size: int):
    if status == _lib.ERR_CODE_SIZE:
        raise EncodingError.CodeSize(detail)
    if status == _lib.ERR_UNEXPECTED_CODE:
        raise EncodingError.UnexpectedCode(detail, code_size)
    if status == _lib.ERR_IO_WRITE_ZERO:
        raise EncodingError("Io", status_message(False, status, detail), io_kind="WriteZero")
    raise EncodingError("Panic", status_message(False, status, detail))


def _encode(params, data, into, code_size):
    status, detail, out = default_codec().encode(params, _read_all(data))
    _write(into, out)  # bytes written before an error stay in the writer, like the reference
    if status != _lib.OK:
        _raise(status, detail, code_size)


class VariableEncoder:
    """encoder.rs:153-347"""

    @staticmethod
    def encode(data, into, code_size: int, endianness: Endianness,
               code_size_strategy: CodeSizeStrategy) -> None:
        """encoder.rs:199-220"""
        _encode(variable_params(code_size, endianness, code_size_strategy), data, into, code_size)

    @staticmethod
    def encode_to_vec(data, code_size: int, endianness: Endianness,
                      code_size_strategy: CodeSizeStrategy) -> bytes:
        """encoder.rs:262-271"""
        output = bytearray()
        VariableEncoder.encode(data, output, code_size, endianness, code_size_strategy)
        return bytes(output)


class GifStyleEncoder:
    """encoder.rs:353-440: little endian, default strategy."""

    @staticmethod
    def encode(data, into, code_size: int) -> None:
        _encode(gif_params(code_size), data, into, code_size)

    @staticmethod
    def encode_to_vec(data, code_size: int) -> bytes:
        output = bytearray()
        GifStyleEncoder.encode(data, output, code_size)
        return bytes(output)


class TiffStyleEncoder:
    """encoder.rs:446-524: code size 8, big endian, TIFF early change."""

    @staticmethod
    def encode(data, into) -> None:
        _encode(tiff_params(), data, into, 8)

    @staticmethod
    def encode_to_vec(data) -> bytes:
        output = bytearray()
        TiffStyleEncoder.encode(data, output)
        return bytes(output)


class FixedEncoder:
    """encoder.rs:530-659: fixed 12-bit codes, no clear / end-of-information codes."""

    @staticmethod
    def encode(data, into, endianness: Endianness) -> None:
        _encode(fixed_params(endianness), data, into, 8)

    @staticmethod
    def encode_to_vec(data, endianness: Endianness) -> bytes:
        output = bytearray()
        FixedEncoder.encode(data, output, endianness)
        return bytes(output)
