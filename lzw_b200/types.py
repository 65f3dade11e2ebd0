"""Option enums of the reference (lzw/src/lib.rs:55-91)."""
import enum

from . import _lib


class Endianness(enum.Enum):
    """lib.rs:59-65: bit ordering of the packed codes."""
    BigEndian = 1      # most significant bit first (TIFF)
    LittleEndian = 0   # least significant bit first (GIF)


class CodeSizeStrategy(enum.Enum):
    """lib.rs:71-91: when the code width grows."""
    Default = 0  # when the dictionary size reaches 2^width        (increment() == 0)
    Tiff = 1     # when the dictionary size reaches 2^width - 1    (increment() == 1)

    def increment(self) -> int:
        return self.value


def variable_params(code_size: int, endianness: Endianness, strategy: CodeSizeStrategy) -> _lib.Params:
    return _lib.Params(_lib.FLAVOUR_VARIABLE, code_size & 0xFF, endianness.value, strategy.value)


def gif_params(code_size: int) -> _lib.Params:
    """encoder.rs:392-399 / decoder.rs:333-340"""
    return variable_params(code_size, Endianness.LittleEndian, CodeSizeStrategy.Default)


def tiff_params() -> _lib.Params:
    """encoder.rs:479-487 / decoder.rs:420-428"""
    return variable_params(8, Endianness.BigEndian, CodeSizeStrategy.Tiff)


def fixed_params(endianness: Endianness) -> _lib.Params:
    """encoder.rs:565-576 / decoder.rs:503-514"""
    return _lib.Params(_lib.FLAVOUR_FIXED, 0, endianness.value, 0)


def lenient_params(code_size: int, endianness: Endianness, strategy: CodeSizeStrategy) -> _lib.Params:
    """Decoder extension (include/slzw.h, SLZW_FLAVOUR_VARIABLE_LENIENT): like variable_params, but a
    full dictionary freezes until the next clear code instead of raising MissingClearCode."""
    return _lib.Params(_lib.FLAVOUR_VARIABLE_LENIENT, code_size & 0xFF, endianness.value, strategy.value)
