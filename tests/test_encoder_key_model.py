"""The algebra the encoder's one-reduction lookup rests on (encode_kernels.cu), checked
exhaustively on the CPU with the constants read from the kernel source:

    q(code) = code * kScr + kScrAdd mod 4096            (codes are kept as q)
    key(prefix_q, byte) = prefix_q * 4096 + (byte << 24 | q(byte))      one IMAD
    slot(key, entry_q)  = ~(key ^ entry_q), empty slot = 0
    hit test            = min over the bucket of ~(slot ^ key) < 4095, and that minimum is entry_q
"""
import os
import re

import numpy as np

from tests.conftest import ROOT

SRC = os.path.join(ROOT, "lzw_b200", "csrc", "encode_kernels.cu")
M32 = np.uint64(0xFFFFFFFF)


def _constants():
    text = open(SRC).read()
    scr = int(re.search(r"constexpr uint32_t kScr = (0x[0-9A-Fa-f]+)u;", text).group(1), 16)
    inv = int(re.search(r"constexpr uint32_t kScrInv = (0x[0-9A-Fa-f]+)u;", text).group(1), 16)
    add = (0x1000 - ((255 * scr) & 0xFFF)) & 0xFFF          # kScrAdd as the source defines it
    assert "kScrAdd = (0x1000u - ((255u * kScr) & 0xFFFu)) & 0xFFFu" in text
    unadd = (0x1000 - ((add * inv) & 0xFFF)) & 0xFFF
    assert "kUnscrAdd = (0x1000u - ((kScrAdd * kScrInv) & 0xFFFu)) & 0xFFFu" in text
    return scr, inv, add, unadd


def test_scramble_is_a_bijection_with_the_two_special_values():
    scr, inv, add, unadd = _constants()
    codes = np.arange(4096, dtype=np.uint64)
    q = (codes * scr + add) & 0xFFF
    assert np.unique(q).size == 4096
    assert np.array_equal((q * inv + unadd) & 0xFFF, codes)      # unscrq inverts scrq
    # q == 4095 must belong to a code that is never a dictionary value: a root for EVERY code size
    assert int(codes[q == 4095][0]) < 4
    # the key whose upper 20 bits are all ones is (prefix q = 4095, byte 255): q(255) must be 0
    assert int(q[255]) == 0
    # a width tag in bits 12..15 of a buffered code does not disturb the low 12 bits of unscrq
    tagged = q | (np.uint64(11) << np.uint64(12))
    assert np.array_equal((tagged * inv + unadd) & 0xFFF, codes)


def test_hit_test_over_all_keys():
    scr, inv, add, unadd = _constants()
    q_of_byte = ((np.arange(256, dtype=np.uint64) * scr + add) & 0xFFF)
    prefix = np.arange(4096, dtype=np.uint64)
    byte = np.arange(256, dtype=np.uint64)
    # all 2^20 keys, as the kernel builds them (32-bit wrap-around of the IMAD)
    key = ((prefix[:, None] * 4096) + ((byte[None, :] << np.uint64(24)) | q_of_byte[None, :])) & M32
    assert np.unique(key).size == 4096 * 256                      # injective
    # a prefix taken from a record after a miss still carries byte' << 24: the product drops it
    dirty = (prefix[:, None] | (np.uint64(0xAB) << np.uint64(24)))
    key_dirty = ((dirty * 4096) + ((byte[None, :] << np.uint64(24)) | q_of_byte[None, :])) & M32
    assert np.array_equal(key, key_dirty)
    # empty slot (0): ~(0 ^ key) = ~key must never look like a hit
    not_key = (~key) & M32
    assert int(not_key.min()) >= 4095
    # ... and exactly one key gets as low as 4095: byte 255 behind the prefix with q = 4095
    assert int((not_key == 4095).sum()) == 1 and int(not_key[4095, 255]) == 4095
    # an occupied slot is never 0, and reads back its entry for its own key
    rng = np.random.default_rng(1)
    entry_q = rng.integers(0, 4095, size=key.shape, dtype=np.uint64)    # 4095 is never an entry
    slot = (~(key ^ entry_q)) & M32
    assert int(slot.min()) > 0
    assert np.array_equal((~(slot ^ key)) & M32, entry_q)
    # any OTHER key compared with that slot gives at least 4096 (sampled: 2^20 x 64 pairs)
    flat_key, flat_slot = key.reshape(-1), slot.reshape(-1)
    for _ in range(64):
        other = flat_key[rng.permutation(flat_key.size)]
        x = (~(flat_slot ^ other)) & M32
        same = other == flat_key
        assert int(x[~same].min()) >= 4096
