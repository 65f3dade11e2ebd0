"""Host-side mirror of lzw/src/decoder.rs: the four decoder types and DecodingError."""
from __future__ import annotations

from . import _lib
from .codec import default_codec, status_message
from .encoder import _read_all, _write
from .types import (CodeSizeStrategy, Endianness, fixed_params, gif_params, tiff_params,
                    variable_params)


class DecodingError(Exception):
    """decoder.rs:15-25.  `kind` is 'Io', 'CodeSize', 'UnexpectedCode' or 'MissingClearCode'."""

    def __init__(self, kind: str, message: str, code: int | None = None,
                 io_kind: str | None = None):
        super().__init__(message)
        self.kind, self.code, self.io_kind = kind, code, io_kind

    @staticmethod
    def CodeSize(code_size: int) -> "DecodingError":
        return DecodingError("CodeSize", status_message(True, _lib.ERR_CODE_SIZE, code_size),
                             code=code_size)

    @staticmethod
    def UnexpectedCode(code: int) -> "DecodingError":
        return DecodingError("UnexpectedCode",
                             status_message(True, _lib.ERR_UNEXPECTED_CODE, code), code=code)

    @staticmethod
    def MissingClearCode() -> "DecodingError":
        return DecodingError("MissingClearCode",
                             status_message(True, _lib.ERR_MISSING_CLEAR_CODE, 0))


def _raise(status: int, detail: int):
    if status == _lib.ERR_CODE_SIZE:
        raise DecodingError.CodeSize(detail)
    if status == _lib.ERR_UNEXPECTED_CODE:
        raise DecodingError.UnexpectedCode(detail)
    if status == _lib.ERR_MISSING_CLEAR_CODE:
        raise DecodingError.MissingClearCode()
    if status == _lib.ERR_IO_UNEXPECTED_EOF:
        raise DecodingError("Io", status_message(True, status, detail), io_kind="UnexpectedEof")
    if status == _lib.ERR_IO_WRITE_ZERO:
        raise DecodingError("Io", status_message(True, status, detail), io_kind="WriteZero")
    raise DecodingError("Panic", status_message(True, status, detail))


def _decode(params, data, into):
    status, detail, out = default_codec().decode(params, _read_all(data))
    _write(into, out)  # bytes decoded before an error stay in the writer, like the reference
    if status != _lib.OK:
        _raise(status, detail)


class VariableDecoder:
    """decoder.rs:54-291"""

    @staticmethod
    def decode(data, into, code_size: int, endianness: Endianness,
               code_size_strategy: CodeSizeStrategy) -> None:
        """decoder.rs:99-120"""
        _decode(variable_params(code_size, endianness, code_size_strategy), data, into)

    @staticmethod
    def decode_to_vec(data, code_size: int, endianness: Endianness,
                      code_size_strategy: CodeSizeStrategy) -> bytes:
        """decoder.rs:163-172"""
        output = bytearray()
        VariableDecoder.decode(data, output, code_size, endianness, code_size_strategy)
        return bytes(output)


class GifStyleDecoder:
    """decoder.rs:296-383"""

    @staticmethod
    def decode(data, into, code_size: int) -> None:
        _decode(gif_params(code_size), data, into)

    @staticmethod
    def decode_to_vec(data, code_size: int) -> bytes:
        output = bytearray()
        GifStyleDecoder.decode(data, output, code_size)
        return bytes(output)


class TiffStyleDecoder:
    """decoder.rs:389-465"""

    @staticmethod
    def decode(data, into) -> None:
        _decode(tiff_params(), data, into)

    @staticmethod
    def decode_to_vec(data) -> bytes:
        output = bytearray()
        TiffStyleDecoder.decode(data, output)
        return bytes(output)


class FixedDecoder:
    """decoder.rs:471-643"""

    @staticmethod
    def decode(data, into, endianness: Endianness) -> None:
        _decode(fixed_params(endianness), data, into)

    @staticmethod
    def decode_to_vec(data, endianness: Endianness) -> bytes:
        output = bytearray()
        FixedDecoder.decode(data, output, endianness)
        return bytes(output)
