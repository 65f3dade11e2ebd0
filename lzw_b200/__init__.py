"""lzw_b200 -- B200-native batched LZW codec, drop-in for the salzweg hot path.

Host-side mirror of the reference's public interface (/root/reference/lzw/src/lib.rs:51-91):
`Endianness`, `CodeSizeStrategy`, modules `encoder` and `decoder`, plus the batched entry point
(`Codec`) the reference lacks.  Everything computes on the GPU through libslzw.so.
"""
from .types import CodeSizeStrategy, Endianness  # noqa: F401
from .codec import Codec, MultiCodec, PinnedBuffer, default_codec, partition_streams  # noqa: F401
from . import decoder, encoder  # noqa: F401

__all__ = ["Endianness", "CodeSizeStrategy", "Codec", "MultiCodec", "partition_streams", "PinnedBuffer", "default_codec", "encoder", "decoder"]
