"""Pins the CPU oracle against every known-answer vector the reference ships.

Citations are /root/reference/lzw/src/<file>:<lines>.  The vectors below are the reference's
own test data (inputs and expected outputs), restated as Python literals.
"""
import hashlib

import pytest

from oracle import oracle as O

D40 = bytes([1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 2, 2, 2, 2,
             2, 1, 1, 1, 0, 0, 0, 0, 2, 2, 2])  # encoder.rs:667-670
D40_GIF2 = bytes([0x8C, 0x2D, 0x99, 0x87, 0x2A, 0x1C, 0xDC, 0x33, 0xA0, 0x02, 0x55, 0x00])  # encoder.rs:682-685
D40_FIXED_LE = bytes([0x1, 0x0, 0x10, 0x0, 0x21, 0x0, 0x3, 0x31, 0x10, 0x1, 0x21, 0x10, 0x4, 0x21, 0x0, 0x6,
                      0x11, 0x0, 0x8, 0x91, 0x10, 0x0, 0x1, 0x0, 0xF, 0x1, 0x0, 0x4, 0x1])  # encoder.rs:808-811
FEW = bytes([0, 0, 1, 3])


# ---- io.rs:334-572 bit I/O vectors -------------------------------------------------------------
def test_bitio_little_endian_vectors():
    assert O.bitwrite(False, [1], [1]) == bytes([0x01])                       # io.rs:428-431
    assert O.bitwrite(False, [4, 1, 6, 6, 2], [3, 3, 3, 3, 4]) == bytes([0x8C, 0x2D])  # io.rs:434-448
    assert O.bitwrite(False, [0xFFF], [12]) == bytes([0xFF, 0x0F])            # io.rs:451-461
    assert O.bitwrite(False, [0xFFFA], [16]) == bytes([0xFA, 0xFF])           # io.rs:464-474
    assert O.bitread(False, [0x01], [1]) == [1]                               # io.rs:334-341
    assert O.bitread(False, [0x8C, 0x2D], [3, 3, 3, 3, 4]) == [4, 1, 6, 6, 2]  # io.rs:344-357
    assert O.bitread(False, [0xFF, 0x0F], [12]) == [0xFFF]                    # io.rs:360-366
    assert O.bitread(False, [0xFA, 0xFF], [16]) == [0xFFFA]                   # io.rs:369-375


def test_bitio_big_endian_vectors():
    assert O.bitwrite(True, [1], [1]) == bytes([0x80])                        # io.rs:477-487
    assert O.bitwrite(True, [4, 1, 6, 6, 2], [3, 3, 3, 3, 4]) == bytes([0x87, 0x62])   # io.rs:490-504
    assert O.bitwrite(True, [0xFFF], [12]) == bytes([0xFF, 0xF0])             # io.rs:507-517
    assert O.bitwrite(True, [0xFFFA], [16]) == bytes([0xFF, 0xFA])            # io.rs:520-531
    assert O.bitread(True, [0x80], [1]) == [1]                                # io.rs:378-384
    assert O.bitread(True, [0x87, 0x62], [3, 3, 3, 3, 4]) == [4, 1, 6, 6, 2]  # io.rs:387-400
    assert O.bitread(True, [0xFF, 0xF0], [12]) == [0xFFF]                     # io.rs:403-409
    assert O.bitread(True, [0xFF, 0xFA], [16]) == [0xFFFA]                    # io.rs:412-418


@pytest.mark.parametrize("big", [False, True])
def test_bitio_iterator_round_trip(big):                                      # io.rs:534-571
    packed = O.bitwrite(big, [0, 1, 0, 2], [12] * 4)
    assert len(packed) == 6
    # the iterator stops at the first short read; a 5th code is not there
    assert O.bitread(big, packed, [12] * 5) == [0, 1, 0, 2]


# ---- encoder.rs:665-835 ------------------------------------------------------------------------
def test_encode_4color_data():                                                # encoder.rs:666-686
    st, _, out = O.encode(O.variable(2, False, False), D40)
    assert st == O.OK and out == D40_GIF2


def test_encode_few_bytes_all_flavours():
    assert O.encode(O.variable(2, False, False), FEW) == (O.OK, 0, bytes([0x04, 0x32, 0x05]))  # :689-702
    assert O.encode(O.gif(2), FEW) == (O.OK, 0, bytes([0x04, 0x32, 0x05]))                     # :705-712
    assert O.encode(O.fixed(False), FEW) == (O.OK, 0, bytes([0, 0, 0, 1, 0x30, 0]))            # :816-824
    assert O.encode(O.tiff(), FEW) == (O.OK, 0, bytes([0x80, 0, 0, 0, 0x10, 0x1C, 0x04]))      # :827-835


def test_encode_is_repeatable():                                              # encoder.rs:715-737
    assert O.encode(O.variable(2, False, False), D40) == O.encode(O.variable(2, False, False), D40)


def test_encode_lorem_ipsum_matches_shipped_golden(lorem, lorem_encoded):     # encoder.rs:740-755
    st, _, out = O.encode(O.variable(7, False, False), lorem)
    assert st == O.OK
    assert out == lorem_encoded
    assert hashlib.sha256(out).hexdigest().startswith("5c3892f19acb2d6f")


def test_encode_unsupported_code_size():                                      # encoder.rs:758-774
    st, detail, out = O.encode(O.variable(10, False, False), bytes([0]))
    assert (st, detail, out) == (O.ERR_CODE_SIZE, 10, b"")


def test_encode_wrong_data_for_code_size():                                   # encoder.rs:777-795
    st, detail, _ = O.encode(O.variable(2, True, False), bytes([0, 1, 8, 3]))
    assert (st, detail) == (O.ERR_UNEXPECTED_CODE, 8)


def test_encode_4color_data_fixed():                                          # encoder.rs:798-813
    st, _, out = O.encode(O.fixed(False), D40)
    assert st == O.OK and out == D40_FIXED_LE


# ---- decoder.rs:649-769 ------------------------------------------------------------------------
def test_decode_4color_data():                                                # decoder.rs:650-672
    assert O.decode(O.variable(2, False, False), D40_GIF2) == (O.OK, 0, D40)


def test_decode_is_repeatable():                                              # decoder.rs:675-700
    assert O.decode(O.variable(2, False, False), D40_GIF2) == O.decode(O.variable(2, False, False), D40_GIF2)


def test_decode_lorem_ipsum(lorem, lorem_encoded):                            # decoder.rs:703-718
    st, _, out = O.decode(O.variable(7, False, False), lorem_encoded)
    assert st == O.OK and out == lorem


def test_decode_unsupported_code_size():                                      # decoder.rs:721-737
    st, detail, out = O.decode(O.variable(10, False, False), bytes([0]))
    assert (st, detail, out) == (O.ERR_CODE_SIZE, 10, b"")


def test_decode_4color_data_fixed():                                          # decoder.rs:740-756
    assert O.decode(O.fixed(False), D40_FIXED_LE) == (O.OK, 0, D40)


def test_decode_bad_data_tiff():                                              # decoder.rs:759-769
    bad = bytes([0x1F, 0x40, 0x3A, 0x00, 0x00, 0x00, 0x44, 0x00, 0x00, 0x44, 0x00, 0x60, 0x54])
    st, detail, _ = O.decode(O.tiff(), bad)
    assert (st, detail) == (O.ERR_UNEXPECTED_CODE, 258)


# ---- doctests: the [0,0,1,3] vectors through every facade preset (lib.rs:13-31 etc.) -----------
@pytest.mark.parametrize("params,packed", [
    (O.gif(2), bytes([0x04, 0x32, 0x05])),
    (O.variable(2, False, False), bytes([0x04, 0x32, 0x05])),
    (O.tiff(), bytes([0x80, 0, 0, 0, 0x10, 0x1C, 0x04])),
    (O.fixed(False), bytes([0, 0, 0, 1, 0x30, 0])),
])
def test_doctest_vectors_round_trip(params, packed):
    assert O.encode(params, FEW) == (O.OK, 0, packed)
    assert O.decode(params, packed) == (O.OK, 0, FEW)


# ---- exploration/src/lib.rs:553-606: raw code sequence of the 40-byte vector -------------------
def test_exploration_raw_codes():
    # fixed flavour emits the raw trie codes as 12-bit values with 256 roots; the exploration
    # crate numbers new entries from 4 (4 roots), so map entries >= 256 back: c - 256 + 4.
    st, _, out = O.encode(O.fixed(False), D40)
    codes = O.bitread(False, out, [12] * 19)
    mapped = [c if c < 256 else c - 256 + 4 for c in codes]
    assert mapped == [1, 4, 4, 2, 7, 7, 5, 6, 8, 2, 10, 1, 12, 13, 4, 0, 19, 0, 8]
