// encode_kernels.cu -- batched LZW encoder for sm_100a.
//
// One warp per stream, one persistent CTA of 28 warps per SM, streams handed out through a global
// work queue in the order chosen by the scheduler.  What limits the number of streams in flight is
// room for their dictionaries (the reference's arena trie, encoder.rs:58-149, becomes an
// open-addressing hash dictionary keyed on (prefix code, next byte) -> code, one u32 per slot
// [prefix':12 | byte:8 | code':12], 4096 slots = 16 KB):
//   * warps 16..27 keep theirs in shared memory (12 x 16 KB, 16 KB-aligned in the shared window so
//     that `base | offset` needs no add);
//   * warps 0..15 keep theirs in TENSOR MEMORY, which a codec otherwise never touches: 16 x 128
//     columns of the SM's 512, accessed with tcgen05.ld / tcgen05.st (32x32b shapes).
// x' = x * kScr + kScrAdd mod 4096 is a bijection of the 12-bit codes ("scrambled" codes): dense
// sequential codes would form runs under an XOR hash, scrambled ones do not, and prefix' ^ hash(byte)
// then collides no more often than a 32-bit multiplicative hash of the whole key (measured,
// profiles/r01_encode_notes.md; the constants of round 2: tools/exp/bucket_sim.c).  Numbering of new entries is insertion order and lookups are
// exact (probing never gives up), so the emitted codes equal the reference's.
//
// Per stream:
//   * the match loop (encoder.rs:313-337) is a dependent chain, one dictionary lookup per input
//     byte, executed by all 32 lanes with warp-uniform control flow.  Default: bucket lookups
//     (match_tile_bucket: one load, one ballot, one shuffle per byte); alternative for shared
//     memory: scalar probes with a speculative next probe and a 32-wide collision window
//     (match_tile);
//   * everything about an input byte that does not depend on the chain (key bits, hash bits,
//     dictionary base) is precomputed by the whole warp into a 64-bit record per byte; the input
//     never sits in shared memory as raw bytes, and the next tile's words are prefetched into
//     registers while the current tile is matched;
//   * emitted codes are buffered as [width:4 | code':12] and bit-packed by the whole warp
//     (LSB-first like io.rs:234-248 or MSB-first like io.rs:296-311) into a shared-memory word
//     window that is written to the output slot with aligned 32-bit stores;
//   * the dictionary reset (encoder.rs:329-333) is a cooperative vectorised clear.
// Semantics follow VariableEncoder::inner_encode (encoder.rs:273-346) and
// FixedEncoder::inner_encode (encoder.rs:618-658) exactly, including the unchecked first byte
// (encoder.rs:311) and the `&mut [u8]`-writer behaviour when the slot is too small.
#include "slzw_device.cuh"

#ifndef SLZW_U0
#define SLZW_U0 4
#endif
#ifndef SLZW_INSERT_SYNC
#define SLZW_INSERT_SYNC 1  // __syncwarp() between a shared-memory insert and the next bucket load
#endif
#ifndef SLZW_TILE
#define SLZW_TILE 96   // input bytes per tile of the default configuration
#endif
#ifndef SLZW_SWARPS
#define SLZW_SWARPS 12  // its warps with a shared-memory dictionary
#endif
#ifndef SLZW_HIT_REDUX
#define SLZW_HIT_REDUX 1  // hit detection: one warp min-reduction (1) or ballot + find-first + shuffle (0)
#endif


namespace slzw {

namespace {

constexpr int kSlots = 4096;
constexpr uint32_t kBucketMask = 127u << 7;                   // byte offset of a 32-slot bucket
constexpr uint32_t kScr = 0xC55u;     // code -> q = code * kScr + kScrAdd mod 4096 (odd => bijection)
constexpr uint32_t kScrInv = 0xFDu;   // kScr * kScrInv == 1 mod 4096
constexpr uint32_t kScrAdd = (0x1000u - ((255u * kScr) & 0xFFFu)) & 0xFFFu;
constexpr uint32_t kUnscrAdd = (0x1000u - ((kScrAdd * kScrInv) & 0xFFFu)) & 0xFFFu;
constexpr uint32_t kByteMul = 0x83Fu;     // byte -> bucket hash (bucket lookups)
constexpr uint32_t kByteMulLat = 0x6A7u;  // byte -> bucket hash (latency variant)
static_assert(((kScr * kScrInv) & 0xFFFu) == 1u, "kScrInv must invert kScr mod 4096");
// the two properties the bucket lookups rest on (below)
static_assert(((2u * kScr + kScrAdd) & 0xFFFu) == 0xFFFu, "q(2) must be 4095");
static_assert(((255u * kScr + kScrAdd) & 0xFFFu) == 0u, "q(255) must be 0");

// The bucket lookups (match_tile_bucket) keep codes as q = code * kScr + kScrAdd mod 4096.  A
// byte's record carries x = byte << 24 | q(byte), the prefix is a q, and the key of (prefix, byte) is
//     k = prefix * 4096 + x  =  byte << 24 | prefix << 12 | q(byte)              -- ONE IMAD
// (the low 12 bits are a function of the byte, so they take no part in telling two keys apart; a
// prefix that still carries the byte << 24 of the record it was taken from loses it in the
// multiplication).  The slot of an entry is ~(k ^ q(entry)); an empty slot is 0.  Then ~(slot ^ k)
//   * is q(entry) for the slot that holds the key -- never 4095, which is q(2), and code 2 is a
//     root for every code size, never a dictionary value;
//   * has a bit above bit 11 set for a slot that holds another key;
//   * is ~k for an empty slot: at least 4096 unless the key is byte 255 after the prefix with
//     q = 4095, and exactly 4095 then, because q(255) = 0;
// so ONE warp reduction, min over the 32 slots of ~(slot ^ k), answers "found?" (below 4095) and
// IS the new prefix -- no compare + ballot + find-first + shuffle (profiles/r02_encode_notes.md).
// An occupied slot is never 0 (it would need q(entry) = 4095).
template <bool Q>
__device__ __forceinline__ uint32_t scrq(uint32_t code) { return (code * kScr + (Q ? kScrAdd : 0u)) & 0xFFFu; }
template <bool Q>
__device__ __forceinline__ uint32_t unscrq(uint32_t e) { return (e * kScrInv + (Q ? kUnscrAdd : 0u)) & 0xFFFu; }
static_assert(((((7u * kScr + kScrAdd) & 0xFFFu) * kScrInv + kUnscrAdd) & 0xFFFu) == 7u, "unscrq must invert scrq");

// Dictionary accesses by 32-bit shared-window address.  They are volatile asm statements so that
// the compiler keeps them exactly where the match loop puts them (in particular the speculative
// probe stays ahead of the branch it speculates on) and in order with each other.
// One bucket (32 slots = 8 rows of 16 bytes) as a warp-collective matrix load: lanes 0..7 pass the
// row addresses, lane i receives row i / 4, bytes 4 * (i % 4) .. + 3 -- slot i of the bucket, what
// a per-lane 32-bit load would return.  The instruction is .aligned, which tells the compiler that
// the warp is converged around it: the ballots and shuffles of the step then need no divergence
// guard.
__device__ __forceinline__ uint32_t tbl_ld_bucket(uint32_t row_addr) {
    uint32_t v;
    asm volatile("ldmatrix.sync.aligned.m8n8.x1.shared.b16 {%0}, [%1];\n" : "=r"(v) : "r"(row_addr));
    return v;
}
__device__ __forceinline__ void tbl_st(uint32_t saddr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;\n" ::"r"(saddr), "r"(v));
}
// code buffer store by shared-window address (no index arithmetic in the loop)
__device__ __forceinline__ void sts_u16(uint32_t saddr, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;\n" ::"r"(saddr), "h"((uint16_t)v) : "memory");
}

// Per-warp working set besides the dictionary.
template <int TILE>
struct EncMisc {
    static constexpr int kRecs = TILE + 8;                      // + lookahead padding
    static constexpr int kCodeBuf = TILE + 16;                  // codes one tile can emit
    static constexpr int kOutWords = (kCodeBuf * 12) / 32 + 4;  // packed window
    alignas(16) uint2 rec[kRecs];  // per input byte: {byte << 24 | q(byte), hash bits | table base} (latency variant: byte << 12)
    uint32_t outw[kOutWords];
    uint16_t codes[kCodeBuf];  // [width:4 | code':12]
};

// Writes one 32-bit word of the packed window to the output slot.  Word `gw` covers stream
// bytes [4*gw - mis, 4*gw - mis + 4); bytes outside [0, lim) are not written.
__device__ __forceinline__ void store_word(uint8_t* dst, uint32_t mis, uint64_t lim, uint64_t gw,
                                           uint32_t v, bool big) {
    if (big) v = __byte_perm(v, 0, 0x0123);
    const int64_t b0 = (int64_t)(gw * 4) - (int64_t)mis;
    if (b0 >= 0 && (uint64_t)b0 + 4 <= lim) {
        *reinterpret_cast<uint32_t*>(dst + b0) = v;
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int64_t b = b0 + j;
            if (b >= 0 && (uint64_t)b < lim) dst[b] = (uint8_t)(v >> (8 * j));
        }
    }
}

// Match state (encoder.rs:289-311), identical in every lane, kept in registers across tiles.
struct MatchState {
    uint32_t t;       // current_prefix' << 20 (scrambled prefix in key position)
    uint32_t ncs;     // tree.len()' (scrambled next code)
    uint32_t ws;      // write_size, encoder.rs:289
    uint32_t mask;    // size_increase_mask, encoder.rs:292
    uint32_t until;   // inserts left until the next width bump / reset (variable) or until the
                      // table is full (fixed)
    uint32_t ncodes;  // codes buffered for the packer
};

// shared-memory address of this lane's row of bucket (prefix ^ hash7) & 127: one LOP3 + one IMAD
// (written out because the compiler otherwise shifts first and masks afterwards: four instructions)
__device__ __forceinline__ uint32_t bucket_row_addr(uint32_t row0, uint32_t prefix, uint32_t h7) {
    uint32_t x, a;
    asm("lop3.b32 %0, %1, %2, 0x7F, 0x28;\n" : "=r"(x) : "r"(prefix), "r"(h7));  // (prefix ^ h7) & 0x7F
    asm("mad.lo.u32 %0, %1, 128, %2;\n" : "=r"(a) : "r"(x), "r"(row0));
    return a;
}

// ~(a ^ b) and ~(a ^ (b & 0xFFF)) as ONE LOP3 each (the compiler emits XOR + NOT)
__device__ __forceinline__ uint32_t xnor32(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, 0, 0xC3;\n" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t xnor_low12(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, 0xFFF, 0x87;\n" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

// index of the most significant set bit (FLO)
__device__ __forceinline__ uint32_t bfind(uint32_t v) {
    uint32_t r;
    asm("bfind.u32 %0, %1;\n" : "=r"(r) : "r"(v));
    return r;
}

__device__ __forceinline__ void clear_table(uint32_t* table, int lane, uint32_t fill = 0u) {
#pragma unroll 4
    for (int j = lane; j < kSlots / 4; j += kWarpSize)
        reinterpret_cast<uint4*>(table)[j] = make_uint4(fill, fill, fill, fill);
    // the table accesses of the match loops are volatile asm statements without a memory clobber
    // (a clobber makes the compiler reload the records around every lookup: 81 -> 91 ms at config
    // 3); this keeps the plain stores above on their side of them
    asm volatile("" ::: "memory");
}

// ---- tensor memory as a second dictionary store ------------------------------------------------
// Shared memory holds twelve 16 KB dictionaries per SM and nothing else limits the number of
// streams in flight.  The SM's 256 KB of tensor memory (512 columns x 128 lanes x 32 bit) is idle
// in this kernel, so sixteen more warps keep their dictionary there: warp w owns the 32 lanes of
// its lane quarter (w % 4, the only ones a warp can address) and 128 columns; one column = one
// 32-slot bucket, `tcgen05.ld.32x32b.x1` hands every lane its slot of the bucket, and an insert
// stores the column back with one lane's word replaced (`tcgen05.st`).
__device__ __forceinline__ uint32_t tmem_ld(uint32_t taddr) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n" : "=r"(v) : "r"(taddr));
    return v;
}
// The loaded register may only be read after the wait; tying it to the statement keeps the
// compiler from scheduling a use above it.
__device__ __forceinline__ void tmem_wait_ld(uint32_t& v) {
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" : "+r"(v)::"memory");
}
__device__ __forceinline__ void tmem_st(uint32_t taddr, uint32_t v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};\n" ::"r"(taddr), "r"(v));
}
__device__ __forceinline__ void tmem_wait_st() {
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
}
// zeroes the 128 columns of this warp's dictionary
__device__ __forceinline__ void tmem_clear(uint32_t tbase) {
    const uint32_t z = 0u;
#pragma unroll
    for (uint32_t c = 0; c < 128u; c += 16u)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
                     "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};\n" ::"r"(tbase + c),
                     "r"(z));
    tmem_wait_st();
}

// ---- bucket variant of the match loop -----------------------------------------------------------
// With 28 warps per SM the encoder is bound by instruction issue, not by the latency of the
// dependent chain (profiles/r01b_bench_encode_ncu_summary.txt: issue slots 87 % busy), so the
// variant with the fewest instructions per input byte wins, whatever its chain length.  Here the
// dictionary is 128 buckets of 32 slots: a bucket is one 128-byte line of shared memory or one
// column of tensor memory, lane l owns slot l of every bucket, and every lookup is ONE load in
// which each lane reads its slot of the key's bucket, one ballot over "my slot holds this key" and
// one shuffle of the matching slot's code.  A bucket fills from slot 0 upwards; a key that is not
// in its bucket while the bucket still has an empty slot is not in the dictionary (miss: the lane
// that owns the first empty slot inserts it); a full bucket overflows into the next one.  There is
// no separate collision path and no speculation: 1.00 to 1.14 bucket loads per input byte on the
// config-3 strips at the format's load factor of up to 0.94 (profiles/r01_encode_notes.md).
// rec[i] = {byte << 24 | q(byte), hash7(byte)} (shared memory) or {byte << 24 | q(byte), dictionary's
// tensor-memory address | hash7(byte)} (tensor memory); `tl` = shared address of this lane's slot
// in bucket 0, `tb` = tensor-memory address of the dictionary.
// MODE says what a miss may have to do in this tile (chosen per tile by the caller from the
// number of inserts left before the next event):
//   0  every miss inserts and nothing else can happen (no width bump, no reset, table not
//      full): no per-miss bookkeeping, codes are stored without a width tag (the packer uses the
//      tile's width);
//   1  every miss checks for the width bump / reset (variable) or for the table filling up
//      (fixed); codes carry their width;
//   2  fixed flavour, table full: lookups only (encoder.rs:645-647).
template <bool FIXED, bool TMEM, int U, int MODE>
__device__ __forceinline__ void match_tile_bucket(uint32_t* __restrict__ table, const uint32_t tl,
                                                  const uint32_t tb, const uint2* __restrict__ rec,
                                                  uint16_t* __restrict__ codes, const int lane,
                                                  const uint32_t len, MatchState& m,
                                                  const uint32_t cs, const uint32_t inc,
                                                  const uint32_t clear_code,
                                                  const uint32_t first_code) {
    // prefix (q): its low 7 bits choose the bucket, its low 12 bits go into the key and are the
    // code a miss emits; after a miss it is the record's x, byte << 24 | q(byte), as it stands
    uint32_t tp = m.t >> 20;
    uint32_t ncs = m.ncs;                         // low 12 bits count, upper bits are garbage
    uint32_t ws = FIXED ? 12u : m.ws;
    uint32_t wtag = ws << 12;
    uint32_t mask = m.mask;
    uint32_t until = m.until;
    const uint32_t cbase = (uint32_t)__cvta_generic_to_shared(codes);
    uint32_t cpa = cbase + 2u * m.ncodes;         // shared address of the next code
    const uint32_t cpa0 = cpa;
    // a bucket fills from slot 0 upwards and nothing is ever removed, so its empty slots are
    // always the lanes >= some count: the first empty slot is this lane's iff the ballot of empty
    // slots equals "all lanes from mine on"
    const uint32_t lanes_ge = 0xFFFFFFFFu << lane;
    // shared-memory dictionary: a bucket is loaded through its 8 row addresses (lane & 7), an
    // insert goes to this lane's own slot; tlx turns the one address into the other
    const uint32_t tlr = (tl & ~0x7Fu) | (((uint32_t)lane & 7u) << 4);
    const uint32_t tlx = tlr ^ tl;

#define SLZW_STEP_B(RC)                                                                         \
    {                                                                                           \
        /* slots hold ~(key ^ q): an empty slot is 0 and ~(slot ^ key) is q for the slot that   \
           holds the key, at least 4095 for every other one */                                   \
        const uint32_t key = tp * 4096u + (RC).x;                                               \
        uint32_t a = TMEM ? ((tp & 0x7Fu) ^ (RC).y) : bucket_row_addr(tlr, tp, (RC).y);         \
        /* Both bucket loads (tcgen05.ld, ldmatrix) are .sync.aligned: the warp is converged at   \
           every step, the compiler puts no divergence guard in front of the ballots and the     \
           shuffle, and the insert of the previous step (same warp, same memory pipe, program    \
           order) is in place before this step's load. */                                        \
        _Pragma("unroll 1") for (;;) {                                                          \
            uint32_t v;                                                                         \
            if (TMEM) {                                                                         \
                v = tmem_ld(a);                                                                 \
                tmem_wait_ld(v);                                                                \
            } else {                                                                            \
                v = tbl_ld_bucket(a);                                                           \
            }                                                                                   \
            /* q of the entry iff my slot holds the key, at least 4095 otherwise; the minimum   \
               over the bucket (one REDUX) says "found" and is the new prefix */                 \
            if (SLZW_HIT_REDUX) {                                                               \
                const uint32_t c = __reduce_min_sync(kFullMask, xnor32(v, key));                     \
                if (c < 4095u) { /* find_word hit, encoder.rs:319-320 */                        \
                    tp = c;                                                                     \
                    break;                                                                      \
                }                                                                               \
            } else {                                                                            \
                const uint32_t x = xnor32(v, key);                                             \
                const uint32_t bal = __ballot_sync(kFullMask, x < 4095u);                       \
                if (bal) {                                                                      \
                    tp = __shfl_sync(kFullMask, x, (int)bfind(bal));                            \
                    break;                                                                      \
                }                                                                               \
            }                                                                                   \
            const uint32_t em = __ballot_sync(kFullMask, v == 0u);                              \
            if (em == 0u) { /* full bucket: the key may have overflowed into the next one */    \
                a = TMEM ? ((a & ~0x7Fu) | ((a + 1u) & 0x7Fu))                                  \
                         : ((a & ~kBucketMask) | ((a + 128u) & kBucketMask));                   \
                continue;                                                                       \
            }                                                                                   \
            /* miss: encoder.rs:322-324 / 645-649 */                                            \
            /* the prefix being emitted: the low 16 bits of tp hold nothing else */             \
            sts_u16(cpa, MODE == 1 ? (tp | wtag) : tp);                                         \
            cpa += 2u;                                                                          \
            if (MODE == 0 || (MODE == 1 && (!FIXED || until != 0u))) {                          \
                const bool mine = em == lanes_ge; /* I own the first empty slot */              \
                const uint32_t entry = xnor_low12(key, ncs);                                     \
                if (TMEM) {                                                                     \
                    tmem_st(a, mine ? entry : v);                                               \
                    tmem_wait_st();                                                             \
                } else {                                                                        \
                    if (mine) tbl_st(a ^ tlx, entry);                                           \
                    /* the next ldmatrix reads this slot on behalf of another lane: order the   \
                       store before it (memory model; measured cost: r02_encode_notes.md) */    \
                    if (SLZW_INSERT_SYNC) __syncwarp();                                         \
                }                                                                               \
                ncs += kScr;                                                                    \
                if (MODE == 1) {                                                                \
                    until--;                                                                    \
                    if (!FIXED && until == 0u) { /* new index == mask, encoder.rs:326 */        \
                        if (ws < 12u) {          /* encoder.rs:327-328 */                       \
                            ws++;                                                               \
                            wtag = ws << 12;                                                    \
                            const uint32_t nm = (1u << ws) - inc;                               \
                            until = nm - mask;                                                  \
                            mask = nm;                                                          \
                        } else { /* encoder.rs:329-333: clear at 12 bits, dictionary restarts */ \
                            sts_u16(cpa, scrq<true>(clear_code) | (12u << 12));                 \
                            cpa += 2u;                                                          \
                            ws = cs + 1u;                                                       \
                            wtag = ws << 12;                                                    \
                            mask = (1u << ws) - inc;                                            \
                            until = mask - first_code + 1u;                                     \
                            ncs = scrq<true>(first_code);                                       \
                            if (TMEM) {                                                         \
                                tmem_clear(tb);                                                 \
                            } else {                                                            \
                                __syncwarp();                                                   \
                                clear_table(table, lane);                                       \
                                __syncwarp();                                                   \
                            }                                                                   \
                        }                                                                       \
                    }                                                                           \
                }                                                                               \
            }                                                                                   \
            /* prefix = this byte: its q sits in the low bits of the record */                  \
            tp = (RC).x;                                                                        \
            break;                                                                              \
        }                                                                                       \
    }

    uint32_t i = 0;
    if constexpr (U >= 8) {
        while (i + 8u <= len) {
            {
                const uint4 ra = *reinterpret_cast<const uint4*>(rec + i);
                const uint4 rb = *reinterpret_cast<const uint4*>(rec + i + 2);
                const uint2 r0 = make_uint2(ra.x, ra.y), r1 = make_uint2(ra.z, ra.w);
                const uint2 r2 = make_uint2(rb.x, rb.y), r3 = make_uint2(rb.z, rb.w);
                SLZW_STEP_B(r0)
                SLZW_STEP_B(r1)
                SLZW_STEP_B(r2)
                SLZW_STEP_B(r3)
            }
            {
                const uint4 ra = *reinterpret_cast<const uint4*>(rec + i + 4);
                const uint4 rb = *reinterpret_cast<const uint4*>(rec + i + 6);
                const uint2 r0 = make_uint2(ra.x, ra.y), r1 = make_uint2(ra.z, ra.w);
                const uint2 r2 = make_uint2(rb.x, rb.y), r3 = make_uint2(rb.z, rb.w);
                SLZW_STEP_B(r0)
                SLZW_STEP_B(r1)
                SLZW_STEP_B(r2)
                SLZW_STEP_B(r3)
            }
            i += 8u;
        }
    }
    if constexpr (U >= 4) {
        while (i + 4u <= len) {
            // four records with two 128-bit loads (i is a multiple of 4, rec is 16-byte aligned)
            const uint4 ra = *reinterpret_cast<const uint4*>(rec + i);
            const uint4 rb = *reinterpret_cast<const uint4*>(rec + i + 2);
            const uint2 r0 = make_uint2(ra.x, ra.y), r1 = make_uint2(ra.z, ra.w);
            const uint2 r2 = make_uint2(rb.x, rb.y), r3 = make_uint2(rb.z, rb.w);
            SLZW_STEP_B(r0)
            SLZW_STEP_B(r1)
            SLZW_STEP_B(r2)
            SLZW_STEP_B(r3)
            i += 4u;
        }
    }
    if constexpr (U >= 2) {
        while (i + 2u <= len) {
            // two records with one 128-bit load (i is even, rec is 16-byte aligned)
            const uint4 rr = *reinterpret_cast<const uint4*>(rec + i);
            const uint2 r0 = make_uint2(rr.x, rr.y), r1 = make_uint2(rr.z, rr.w);
            SLZW_STEP_B(r0)
            SLZW_STEP_B(r1)
            i += 2u;
        }
    }
#pragma unroll 1
    while (i < len) {
        const uint2 r0 = rec[i];
        SLZW_STEP_B(r0)
        i += 1u;
    }
#undef SLZW_STEP_B

    if (MODE == 0) until -= (cpa - cpa0) >> 1;  // one insert per emitted code
    m.t = tp << 20;
    m.ncs = ncs & 0xFFFu;
    m.ws = ws;
    m.mask = mask;
    m.until = until;
    m.ncodes = (cpa - cbase) >> 1;
}

// ---- latency variant of the match loop: few streams, long streams ---------------------------------
// A batch with no more streams than the device has shared-memory dictionaries (config 4 sharded
// over 8 GPUs: 512 frames of 1 MiB per GPU; a single-stream call) is bound by the dependent chain
// of ONE stream, not by instruction issue.  The bucket lookups above pay a collective load, a warp
// reduction and a ballot per byte (about 270 cycles per byte on a 1 MiB frame); here every lane
// runs the same scalar chain and nothing crosses lanes: the dictionary is 512 buckets of 8 slots,
// a lookup is two 128-bit shared-memory loads of one bucket (a broadcast: all lanes read the same
// address), a min tree over slot ^ ~key (same complemented slots and q codes as above: the minimum
// is below 4095 iff the bucket holds the key, and then it is the new prefix), and a full bucket
// continues at bucket + step(byte) (odd step: double hashing over buckets, 1.18 buckets per byte
// on the config-3 strips, tools/exp/probe_sim.c).  LDS -> 3 ALU levels -> compare -> branch.
// rec[i] = {byte << 12 | q(byte), hash9(byte) << 5}; tb = shared address of the 16 KB table.
template <bool FIXED>
__device__ __forceinline__ void match_tile_lat(uint32_t* __restrict__ table, const uint32_t tb,
                                               const uint2* __restrict__ rec, uint16_t* __restrict__ codes,
                                               const int lane, const uint32_t len, MatchState& m,
                                               const uint32_t cs, const uint32_t inc,
                                               const uint32_t clear_code, const uint32_t first_code) {
    constexpr uint32_t kMask8 = 0x3FE0u;  // byte offset of a 32-byte bucket
    uint32_t q = m.t >> 20;
    uint32_t ncs = m.ncs;
    uint32_t ws = FIXED ? 12u : m.ws;
    uint32_t mask = m.mask;
    uint32_t until = m.until;
    // codes carry a width tag only when a width change can happen inside this tile
    uint32_t wtag = (!FIXED && until <= len) ? ws << 12 : 0u;
    const uint32_t cbase = (uint32_t)__cvta_generic_to_shared(codes);
    uint32_t cpa = cbase + 2u * m.ncodes;
#pragma unroll 2
    for (uint32_t i = 0; i < len; i++) {
        const uint2 rc = rec[i];
        const uint32_t nkey = ~((q << 20) | (rc.x & 0xFF000u));
        uint32_t a = tb | (((q << 5) ^ rc.y) & kMask8);
#pragma unroll 1
        for (;;) {
            uint4 e0, e1;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(e0.x), "=r"(e0.y), "=r"(e0.z), "=r"(e0.w) : "r"(a));
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(e1.x), "=r"(e1.y), "=r"(e1.z), "=r"(e1.w) : "r"(a + 16u));
            const uint32_t m0 = min(min(e0.x ^ nkey, e0.y ^ nkey), e0.z ^ nkey);
            const uint32_t m1 = min(min(e0.w ^ nkey, e1.x ^ nkey), e1.y ^ nkey);
            const uint32_t m2 = min(e1.z ^ nkey, e1.w ^ nkey);
            const uint32_t c = min(min(m0, m1), m2);
            if (c < 4095u) {  // find_word hit, encoder.rs:319-320
                q = c;
                break;
            }
            if (e1.w != 0u) {  // full bucket without the key: it may have gone to the next one
                const uint32_t stp = ((rc.x << 5) & kMask8) | 32u;
                a = (a & ~kMask8) | ((a + stp) & kMask8);
                continue;
            }
            // miss: encoder.rs:322-324 / 645-649
            sts_u16(cpa, q | wtag);
            cpa += 2u;
            if (!FIXED || until != 0u) {
                // first empty slot of the bucket (slots fill from 0 upwards); every lane stores the
                // same word to the same address
                const bool up = e0.w != 0u;
                const uint32_t s0 = up ? e1.x : e0.x, s1 = up ? e1.y : e0.y, s2 = up ? e1.z : e0.z;
                const bool t1 = s1 == 0u;
                const uint32_t lo = t1 ? s0 : s2;
                const uint32_t idx = (up ? 4u : 0u) + (t1 ? 0u : 2u) + (lo != 0u ? 1u : 0u);
                tbl_st(a + 4u * idx, nkey & ~(ncs & 0xFFFu));
                ncs += kScr;
                until--;
                if (!FIXED && until == 0u) {  // new index == mask, encoder.rs:326
                    if (ws < 12u) {           // encoder.rs:327-328
                        ws++;
                        wtag = ws << 12;
                        const uint32_t nm = (1u << ws) - inc;
                        until = nm - mask;
                        mask = nm;
                    } else {                  // encoder.rs:329-333: clear at 12 bits, dictionary restarts
                        sts_u16(cpa, scrq<true>(clear_code) | (12u << 12));
                        cpa += 2u;
                        ws = cs + 1u;
                        wtag = ws << 12;
                        mask = (1u << ws) - inc;
                        until = mask - first_code + 1u;
                        ncs = scrq<true>(first_code);
                        __syncwarp();
                        clear_table(table, lane);
                        __syncwarp();
                    }
                }
            }
            q = rc.x & 0xFFFu;  // prefix = this byte
            break;
        }
    }
    m.t = q << 20;
    m.ncs = ncs & 0xFFFu;
    m.ws = ws;
    m.mask = mask;
    m.until = until;
    m.ncodes = (cpa - cbase) >> 1;
}

// The 32-bit word of lane `lane` of the tile whose first byte is at `p` (skew = p & 3): the
// aligned word at p - skew + 4 * lane, or 0 when it holds no byte of [p, p + tile_len).  A word
// that lies partly outside the tile is assembled from byte loads, so nothing outside
// [p, p + tile_len) is ever touched (the input may be pinned host memory read in place).
// The loads go to L2 (ld.global.cg): every byte is read once, and in a streaming launch
// (StreamCtl) the bytes behind a window boundary are written by the copy engine while the kernel
// runs -- a 32-byte sector that straddles the boundary must not be served from an L1 line that
// was filled before they arrived.
__device__ __forceinline__ uint32_t load_tile_word(const uint8_t* __restrict__ p, uint32_t skew,
                                                   uint32_t tile_len, int lane) {
    const uint32_t lo = 4u * (uint32_t)lane;
    const uint32_t end = skew + tile_len;
    if (lo >= skew && lo + 4u <= end) return __ldcg(reinterpret_cast<const uint32_t*>(p - skew + lo));
    uint32_t v = 0u;
    if (lo + 3u >= skew && lo < end) {
#pragma unroll
        for (uint32_t b = 0; b < 4u; b++)
            if (lo + b >= skew && lo + b < end) v |= (uint32_t)__ldcg(p - skew + lo + b) << (8u * b);
    }
    return v;
}

// tmem (warp-uniform) = the stream's dictionary lives in tensor memory at address `tb` (table is
// unused); otherwise `table` / `tb` are the generic pointer and the shared-window address of its
// 16 KB.  One function body for both kinds of warp: everything but the match loop is shared, which
// keeps the instruction footprint of the 28 warps down.
template <int TILE, bool FIXED, bool HAS_TMEM, int U, bool LAT = false>
__device__ void encode_stream(const DevBatch& a, uint32_t sid, uint32_t* __restrict__ table,
                              const uint32_t tb, const bool tmem_warp, EncMisc<TILE>& S, int lane) {
    const bool TMEM = HAS_TMEM && tmem_warp;
    using Misc = EncMisc<TILE>;
    static_assert(TILE % 4 == 0 && TILE <= 4 * kWarpSize, "one 32-bit word per lane");
    uint2* __restrict__ rec = S.rec;
    constexpr int kRecByteShift = LAT ? 12 : 24;  // where a record's x holds the byte
    uint16_t* __restrict__ codes = S.codes;
    uint32_t* __restrict__ outw = S.outw;
    const uint64_t in_begin = a.in_off[sid];
    const uint64_t n = a.in_off[sid + 1] - in_begin;
    const uint8_t* src = a.in + in_begin;
    uint8_t* dst = nullptr;
    uint64_t cap = ~0ull;
    if (a.out != nullptr) {
        const uint64_t ob = a.out_off[sid];
        dst = a.out + ob;
        cap = a.out_off[sid + 1] - ob;
    }
    const bool big = a.p.big_endian != 0;
    const uint32_t inc = (!FIXED && a.p.tiff_early_change) ? 1u : 0u;
    const uint32_t cs = FIXED ? 8u : (a.code_size ? a.code_size[sid] : a.p.code_size);

    if (!FIXED && (cs < 2 || cs > 8)) {  // encoder.rs:281-283
        if (lane == 0) {
            a.out_len[sid] = 0;
            a.status[sid] = SLZW_ERR_CODE_SIZE;
            a.detail[sid] = cs;
        }
        return;
    }

    // first tile's words in flight while the dictionary is cleared; the first byte of the stream
    // is consumed below (encoder.rs:311), tiles start at byte 1
    uint64_t pos = n > 0 ? 1 : 0;
    uint32_t skew = (uint32_t)(reinterpret_cast<uintptr_t>(src + pos) & 3u);
    uint32_t tile_len = (uint32_t)((n - pos) < (uint64_t)(TILE - skew) ? (n - pos) : (TILE - skew));
    uint32_t w = tile_len ? load_tile_word(src + pos, skew, tile_len, lane) : 0u;

    if (TMEM) tmem_clear(tb);
    else clear_table(table, lane);
    for (int i = lane; i < Misc::kOutWords; i += kWarpSize) outw[i] = 0;

    const uint32_t max_code = (1u << cs) - 1;  // encoder.rs:285
    const uint32_t clear_code = 1u << cs;      // encoder.rs:290
    const uint32_t eoi = clear_code + 1;       // encoder.rs:291
    const uint32_t first_code = FIXED ? 256u : clear_code + 2;

    MatchState m;
    m.t = 0;
    m.ncs = scrq<true>(first_code);
    m.ws = FIXED ? 12u : cs + 1;
    m.mask = (1u << m.ws) - inc;
    m.until = FIXED ? 4096u - first_code : m.mask - first_code + 1u;
    m.ncodes = 0;
    uint32_t status = SLZW_OK, detail = 0;

    // ---- packed-window state (warp-uniform) ----
    const uint32_t mis = dst ? (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 3u) : 0u;
    uint32_t qbits = 8 * mis;  // bit cursor inside the window
    uint64_t wbase = 0;        // aligned words already flushed
    uint64_t bits = 0;         // bits handed to the bit writer so far
    // the writer fails when byte index `cap` is written: (bits >> 3) > cap  <=>  bits > limit
    const uint64_t limit = cap > (~0ull - 7) / 8 ? ~0ull : cap * 8 + 7;

    auto push = [&](uint32_t code, uint32_t width) {  // BitWriter::write, io.rs:234-237, 296-300
        codes[m.ncodes++] = (uint16_t)(scrq<true>(code) | (width << 12));
    };

    // Whole warp: bit-pack the buffered codes, flush complete words.
    // An entry without a width tag (stored by a match loop that does not track the width) takes
    // the width `wdef` in front of buffer index `split` and `wdef + 1` from there on (the one width
    // bump such a tile may contain, see below).
    auto pack_and_flush = [&](uint32_t count, uint32_t wdef, uint32_t split = ~0u) {
        for (uint32_t base = 0; base < count; base += kWarpSize) {
            const uint32_t idx = base + lane;
            const uint32_t e = idx < count ? codes[idx] : 0u;
            const uint32_t w0 = base < split ? wdef : wdef + 1u;  // untagged width at the round's start
            const bool mixed = split > base && split < base + (uint32_t)kWarpSize;  // ... changes inside it
            const uint32_t wu = idx < split ? wdef : wdef + 1u;
            const uint32_t wd = idx < count ? ((e >> 12) ? (e >> 12) : wu) : 0u;
            const uint32_t code = unscrq<true>(e) & ((1u << wd) - 1u);  // BitWriter masks to the width
            // inclusive prefix sum of the widths; codes without a tag all have the same width in
            // nearly every round (MODE 0 / 2 tiles), which needs no scan
            uint32_t x, total;
            if (!mixed && !__any_sync(kFullMask, (e >> 12) != 0u)) {
                const uint32_t left = count - base;
                x = ((uint32_t)lane + 1u) * w0;
                total = (left < (uint32_t)kWarpSize ? left : (uint32_t)kWarpSize) * w0;
            } else {
                x = wd;
#pragma unroll
                for (int d = 1; d < kWarpSize; d <<= 1) {
                    const uint32_t y = __shfl_up_sync(kFullMask, x, d);
                    if (lane >= d) x += y;
                }
                total = __shfl_sync(kFullMask, x, kWarpSize - 1);
            }
            const uint32_t off = qbits + x - wd;
            if (wd) {
                const uint32_t wi = off >> 5, sh = off & 31u;
                if (!big) {
                    const uint64_t v = (uint64_t)code << sh;
                    atomicOr(&outw[wi], (uint32_t)v);
                    if (v >> 32) atomicOr(&outw[wi + 1], (uint32_t)(v >> 32));
                } else {
                    const uint64_t v = (uint64_t)code << (64 - wd - sh);
                    atomicOr(&outw[wi], (uint32_t)(v >> 32));
                    if ((uint32_t)v) atomicOr(&outw[wi + 1], (uint32_t)v);
                }
            }
            qbits += total;
            bits += total;
        }
        __syncwarp();
        const uint32_t cw = qbits >> 5;
        if (dst)
            for (uint32_t k = lane; k < cw; k += kWarpSize)
                store_word(dst, mis, cap, wbase + k, outw[k], big);
        const uint32_t carry = outw[cw];
        __syncwarp();
        for (uint32_t k = lane; k <= cw; k += kWarpSize) outw[k] = (k == 0) ? carry : 0u;
        wbase += cw;
        qbits &= 31u;
        __syncwarp();
    };

    __syncwarp();

    if (!FIXED) push(clear_code, m.ws);  // encoder.rs:297
    if (n > 0) {
        const uint32_t first = __ldcg(src);  // encoder.rs:311 / 637: not range-checked
        m.t = scrq<true>(first) << 20;
        if (!FIXED && n > 1 && first >= first_code) {
            // find_word would index past tree.nodes (encoder.rs:99) unless the second byte
            // is rejected first (encoder.rs:315-317)
            const uint32_t k = __ldcg(src + 1);
            if (k > max_code) {
                status = SLZW_ERR_UNEXPECTED_CODE;
                detail = k;
            } else {
                status = SLZW_ERR_REFERENCE_PANIC;
            }
        }
    }
    // the leading clear code alone overflows a tiny slot before the first byte is even read
    if (!FIXED && (uint64_t)m.ws > limit) {
        status = SLZW_ERR_IO_WRITE_ZERO;
        detail = 0;
    }

    while (status == SLZW_OK && pos < n) {
        // ---- records of this tile; input validation (encoder.rs:315-317) truncates it ----
        uint32_t bad = 0xFFFFFFFFu;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const uint32_t idx = 4u * (uint32_t)lane + (uint32_t)b - skew;  // wraps before the tile
            const uint32_t k = (w >> (8 * b)) & 0xFFu;
            if (idx < tile_len) {
                const uint32_t h7 = ((k * kByteMul) >> 2) & 0x7Fu;
                // the low 12 bits carry q of the byte itself (the prefix after a miss)
                rec[idx] = make_uint2((k << kRecByteShift) | scrq<true>(k), TMEM ? (tb | h7)
                                               : LAT ? (((k * kByteMulLat) << 5) & 0x3FE0u)
                                                     : h7);
                if (!FIXED && k > max_code && idx < bad) bad = idx;
            }
        }
        // the lookahead of the tile's last byte reads one record past the end: keep its probe
        // address inside the dictionary
        if (lane == 0) rec[tile_len] = make_uint2(0u, tb);
        uint32_t len = tile_len;
        if (!FIXED && cs < 8) {
            bad = __reduce_min_sync(kFullMask, bad);
            if (bad < len) len = bad;
        }
        // next tile's words (aligned from here on) in flight during the match loop
        const uint64_t npos = pos + tile_len;
        const uint32_t nlen = (uint32_t)((n - npos) < (uint64_t)TILE ? (n - npos) : TILE);
        const uint32_t wn = nlen ? load_tile_word(src + npos, 0u, nlen, lane) : 0u;
        __syncwarp();

        const uint32_t ws_tile = m.ws;  // width of the codes a MODE 0 / 2 tile stores without a tag
        uint32_t split = ~0u;           // buffer index of the first untagged code of width ws_tile + 1
        if (len) {
            // inserts left before the next event against the most this tile can insert
            int mode = (FIXED && m.until == 0u) ? 2 : (m.until > len ? 0 : 1);
            // A width bump (encoder.rs:327-328) changes nothing but the width of the codes that
            // follow it: a tile that can meet ONE bump and nothing else runs the MODE 0 loop, the
            // packer takes the new width from the until-th code of the tile on, and the state is
            // brought up to date below.  Left to MODE 1: the reset at 12 bits and tiles that could
            // meet two events (small code sizes at the start of a stream).
            bool bump = false;
            if (!FIXED && !LAT && mode == 1 && m.ws < 12u) {
                const uint32_t nm = (2u << m.ws) - inc;  // size_increase_mask after the bump
                if (m.until + (nm - m.mask) > len) {
                    bump = true;
                    split = m.ncodes + m.until;
                    mode = 0;
                }
            }
            // unroll of the MODE 0 and MODE 2 loops: 4 steps per iteration amortise the loop's own
            // instructions (81.9 against 85.0 ms at config 3, 51.7 against 53.7 ms at config 5);
            // 8 spill and run at 90.6 ms.  MODE 1 tiles are rare and keep U.
            constexpr int kU0 = SLZW_U0;
#define SLZW_MATCH_B(TM, MD)                                                                        \
    match_tile_bucket<FIXED, TM, (MD == 1 ? U : kU0), MD>(table, TM ? 0u : (tb | (4u * (uint32_t)lane)), TM ? tb : 0u, rec, \
                                        codes, lane, len, m, cs, inc, clear_code, first_code)
            if constexpr (LAT) {
                match_tile_lat<FIXED>(table, tb, rec, codes, lane, len, m, cs, inc, clear_code, first_code);
            } else if (TMEM) {
                if (mode == 0) SLZW_MATCH_B(true, 0);
                else if (mode == 1) SLZW_MATCH_B(true, 1);
                else if constexpr (FIXED) SLZW_MATCH_B(true, 2);
            } else {
                if (mode == 0) SLZW_MATCH_B(false, 0);
                else if (mode == 1) SLZW_MATCH_B(false, 1);
                else if constexpr (FIXED) SLZW_MATCH_B(false, 2);
            }
#undef SLZW_MATCH_B
            // MODE 0 left until = until - inserts (mod 2^32); the until-th insert was the bump
            if (bump && (int32_t)m.until <= 0) {
                m.ws++;
                const uint32_t nm = (1u << m.ws) - inc;
                m.until += nm - m.mask;
                m.mask = nm;
            }
        }
        __syncwarp();

        pack_and_flush(m.ncodes, ws_tile, split);
        m.ncodes = 0;
        // The `&mut [u8]` writer fails at the first byte past the slot (io.rs:244 / 307).  Codes
        // are emitted in input order, so a write failure inside this tile precedes a rejected
        // byte that ended it.
        if (bits > limit) {
            status = SLZW_ERR_IO_WRITE_ZERO;
        } else if (len < tile_len) {
            status = SLZW_ERR_UNEXPECTED_CODE;
            detail = (rec[len].x >> kRecByteShift) & 0xFFu;
        }
        pos = npos;
        tile_len = nlen;
        skew = 0;
        w = wn;
    }

    // tail codes (only when the whole input was consumed), then pack whatever is buffered --
    // on an error the codes written before it stay in the output, like the reference's writer
    if (status == SLZW_OK) {
        if (n > 0) codes[m.ncodes++] = (uint16_t)((m.t >> 20) | (m.ws << 12));  // encoder.rs:339 / 653
        if (!FIXED) push(eoi, m.ws);                                            // encoder.rs:303 / 340
    }
    __syncwarp();
    pack_and_flush(m.ncodes, m.ws);
    m.ncodes = 0;
    if (status == SLZW_OK && bits > limit) status = SLZW_ERR_IO_WRITE_ZERO;
    const bool finished = (status == SLZW_OK);  // the encoder reached fill()

    if (lane == 0) {
        // fill() (io.rs:251-259, 314-322) only runs when the encoder reached its end
        uint64_t total = finished ? ((bits + 7) >> 3) : (bits >> 3);
        if (finished && total > cap) status = SLZW_ERR_IO_WRITE_ZERO;
        if (total > cap) total = cap;
        if (dst) {
            // bytes of the last, partial window word
            const int64_t b0 = (int64_t)(wbase * 4) - (int64_t)mis;
            uint32_t v = outw[0];
            if (big) v = __byte_perm(v, 0, 0x0123);
            for (int j = 0; j < 4; j++) {
                const int64_t b = b0 + j;
                if (b >= 0 && (uint64_t)b < total) dst[b] = (uint8_t)(v >> (8 * j));
            }
        }
        a.out_len[sid] = total;
        a.status[sid] = status;
        a.detail[sid] = detail;
    }
    __syncwarp();
}

#ifdef SLZW_EXP_LANES
#include "exp/encode_lanes.cuh"
#endif

// Shared-memory layout of the CTA (dynamic shared memory, `base` = shared-window address of its
// first usable byte): the SWARPS dictionaries start at the first 16 KB boundary; the per-warp
// EncMisc blocks of all WARPS warps fill the gap in front of it and continue behind the last
// dictionary.
template <int TILE, int SWARPS, int WARPS>
struct EncLayout {
    static constexpr uint32_t kTable = kSlots * 4;
    static constexpr uint32_t kMisc = (uint32_t)((sizeof(EncMisc<TILE>) + 15) & ~size_t(15));
    static constexpr uint32_t kHead = 16;  // tensor-memory base address slot
    __host__ __device__ static uint32_t first_table(uint32_t base) {
        return (base + kTable - 1) & ~(kTable - 1);
    }
    __host__ __device__ static uint32_t misc_in_front(uint32_t base) {
        const uint32_t k = (first_table(base) - base) / kMisc;
        return k < (uint32_t)WARPS ? k : (uint32_t)WARPS;
    }
    __host__ __device__ static uint32_t misc_addr(uint32_t base, uint32_t warp) {
        const uint32_t front = misc_in_front(base);
        return warp < front ? base + warp * kMisc
                            : first_table(base) + SWARPS * kTable + (warp - front) * kMisc;
    }
    __host__ __device__ static uint32_t bytes(uint32_t base) {  // bytes needed from `base` on
        const uint32_t front = misc_in_front(base);
        return first_table(base) - base + SWARPS * kTable + (WARPS - front) * kMisc;
    }
};

// Warps [0, TWARPS) keep their dictionary in tensor memory, warps [TWARPS, TWARPS + SWARPS) in
// shared memory.
template <int TILE, int SWARPS, int TWARPS, int U, bool FIXED, bool LAT = false>
__global__ void __launch_bounds__((SWARPS + TWARPS) * kWarpSize, 1)
slzw_encode_kernel(const DevBatch a, const uint32_t dyn_bytes) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    constexpr int WARPS = SWARPS + TWARPS;
    using L = EncLayout<TILE, SWARPS, WARPS>;
    static_assert(TWARPS == 0 || TWARPS == 16, "tensor memory holds 16 dictionaries of 128 columns");
    const uint32_t warp = threadIdx.x / kWarpSize;
    const int lane = threadIdx.x % kWarpSize;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem_raw) + L::kHead;
    if (L::bytes(base) + L::kHead > dyn_bytes) __trap();  // launch configuration and layout disagree

    uint32_t tmem_base = 0;
    if constexpr (TWARPS > 0) {
        if (warp == 0) {  // one warp allocates all 512 columns for the CTA (one CTA per SM)
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                             (uint32_t)__cvta_generic_to_shared(smem_raw)),
                         "r"(512u));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_raw);
    }

    EncMisc<TILE>& S =
        *reinterpret_cast<EncMisc<TILE>*>(smem_raw + L::kHead + (L::misc_addr(base, warp) - base));
    const bool tmem_warp = TWARPS > 0 && warp < (uint32_t)(TWARPS > 0 ? TWARPS : 1);
    uint32_t tb;
    uint32_t* table = nullptr;
    if (tmem_warp) {
        // lane quarter warp % 4 (address bits 16 and up), columns 128 * (warp / 4)
        tb = tmem_base + (((warp & 3u) * 32u) << 16) + (warp >> 2) * 128u;
    } else {
        tb = L::first_table(base) + (warp - TWARPS) * L::kTable;
        table = reinterpret_cast<uint32_t*>(smem_raw + L::kHead + (tb - base));
    }
    // A batch smaller than the grid's warps spreads over the SMs instead of filling the first CTAs
    // that come up: only ceil(n / CTAs) warps per CTA take streams (the first ones: consecutive warp
    // ids sit on different schedulers).  Fewer streams per SM means less contention for the issue
    // slots, i.e. a shorter chain per byte for every stream of a small batch.
    const uint32_t takers = (uint32_t)((a.n + gridDim.x - 1) / gridDim.x);
    // (which kind of warp takes first makes no difference: 104.4 ms with the tensor-memory warps
    // first, 103.1 ms with the shared-memory ones, 512 frames of 1 MiB)
    for (; warp < takers;) {
        unsigned long long q = 0;
        if (lane == 0) q = atomicAdd(a.queue, 1ull);
        q = __shfl_sync(kFullMask, q, 0);
        if (q >= a.n) break;
        const uint32_t sid = a.order ? a.order[q] : (uint32_t)q;
        uint32_t win = 0;
        bool arrived = true;
        if (a.sc.win_of) {  // streaming launch: wait for the window of this stream
            win = a.sc.win_of[q];
            const long long t0 = clock64();
            for (;;) {
                uint32_t v;
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(a.sc.avail) : "memory");
                if (v > win) break;
                __nanosleep(500);
                if (clock64() - t0 > (16ll << 30)) {  // about 9 s: the copy never came
                    arrived = false;
                    break;
                }
            }
        }
        if (arrived) {
            encode_stream<TILE, FIXED, (TWARPS > 0), U, LAT>(a, sid, table, tb, tmem_warp, S, lane);
        } else if (lane == 0) {
            a.out_len[sid] = 0;
            a.status[sid] = SLZW_ERR_IO_UNEXPECTED_EOF;
            a.detail[sid] = 0;
            *a.sc.host_abort = 1u;
        }
        if (a.sc.win_of) {
            // everything this warp wrote for the stream is visible before the stream counts as done
            __threadfence();
            __syncwarp();
            if (lane == 0) {
                const uint32_t before = atomicAdd(&a.sc.done[win], 1u);
                if (before + 1u == a.sc.win_count[win]) {
                    __threadfence_system();
                    a.sc.host_flags[win] = 1u;
                }
            }
        }
    }

    if constexpr (TWARPS > 0) {
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
        if (warp == 0)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512u));
    }
}


#ifdef SLZW_EXP_LANES
#include "exp/encode_lanes_kernel.cuh"
#endif

}  // namespace

// ---- launch configuration ---------------------------------------------------------------------
// {input tile, warps with a shared-memory dictionary, warps with a tensor-memory dictionary}
// LAT (latency variant) exists for the fixed flavour only.
template <int TILE, int SWARPS, int TWARPS, int U, bool LAT = false>
struct EncConfig {
    using L = EncLayout<TILE, SWARPS, SWARPS + TWARPS>;
    // the shared window of a CTA starts with 1 KB reserved by the system; taking the larger of
    // the two layouts keeps the launch valid should the dynamic region start at 0 instead
    static uint32_t smem() {
        const uint32_t a = L::bytes(1024u + L::kHead), b = L::bytes(L::kHead);
        return (a > b ? a : b) + L::kHead;
    }
    static cudaError_t configure() {
        if constexpr (!LAT) {
            cudaError_t e = cudaFuncSetAttribute(slzw_encode_kernel<TILE, SWARPS, TWARPS, U, false, LAT>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem());
            if (e != cudaSuccess) return e;
        }
        return cudaFuncSetAttribute(slzw_encode_kernel<TILE, SWARPS, TWARPS, U, true, LAT>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem());
    }
    static cudaError_t launch(const DevBatch& a, int num_sms, cudaStream_t stream) {
        constexpr int WARPS = SWARPS + TWARPS;
        // one CTA per SM as soon as there is a stream for each (the kernel spreads a small batch)
        const int grid = (int)(a.n < (uint64_t)num_sms ? a.n : (uint64_t)num_sms);
        if (a.p.flavour == SLZW_FLAVOUR_FIXED)
            slzw_encode_kernel<TILE, SWARPS, TWARPS, U, true, LAT><<<grid, WARPS * kWarpSize, smem(), stream>>>(a, smem());
        else if constexpr (!LAT)
            slzw_encode_kernel<TILE, SWARPS, TWARPS, U, false, LAT><<<grid, WARPS * kWarpSize, smem(), stream>>>(a, smem());
        return cudaGetLastError();
    }
};

#ifdef SLZW_EXP_LANES
#include "exp/encode_lanes_config.cuh"
#endif

// {input tile, shared-memory dictionaries, tensor-memory dictionaries, step unrolling}: 28 streams
// per SM, one warp each.  Round 1 kept a second configuration (scalar speculative probes, 12 warps)
// for batches with few streams; with the min-reduction lookups this one is faster there too
// (config 4, 512 frames of 1 MiB: 148.7 against 171.9 ms), so every batch takes the same kernel.
using Enc0 = EncConfig<SLZW_TILE, SLZW_SWARPS, 16, 2>;
// Latency variant (match_tile_lat): 12 warps per SM, shared-memory dictionaries only, for batches
// that would leave most of Enc0's 28 warps per SM without a stream -- FIXED flavour only, where it
// wins (1,776 text chunks of 64 KiB: 10.5 against 16.3 ms; once the table is full nothing but
// lookups is left).  On the variable flavours it loses (1 MiB GIF frames: 175 against 148 ms, TIFF
// strips: 11.1 against 8.9 ms: its miss path is longer), so they keep the bucket lookups whatever
// the batch size (profiles/r02_encode_notes.md).
using EncLat = EncConfig<128, 12, 0, 2, true>;

#ifndef SLZW_LANE_CONFIGS
#define SLZW_LANE_CONFIGS(X)  // experiment configurations (exp/encode_lanes_config.cuh) are not built in
#endif
static int g_enc_config = 0;

// tests and tuning (SLZW_ENC_CONFIG): 1 / 2 force the latency (fixed flavour) / throughput variant
// whatever the batch size; 10.. are the configurations of exp/ when they are built in
void encode_select_config(int c) { g_enc_config = c; }

cudaError_t encode_configure() {
    cudaError_t e = Enc0::configure();
    if (e != cudaSuccess) return e;
    if ((e = EncLat::configure()) != cudaSuccess) return e;
#define X(id, C) if ((e = C::configure()) != cudaSuccess) return e;
    SLZW_LANE_CONFIGS(X)
#undef X
    return cudaSuccess;
}

// global-memory dictionaries the launch needs (none outside the experiments)
size_t encode_table_bytes(uint64_t, int num_sms) {
    (void)num_sms;
    switch (g_enc_config) {
#define X(id, C) case id: return C::table_bytes(num_sms);
        SLZW_LANE_CONFIGS(X)
#undef X
        default: return 0;
    }
}

cudaError_t encode_launch(const DevBatch& a, int num_sms, cudaStream_t stream) {
    switch (g_enc_config) {
#define X(id, C) case id: return C::launch(a, num_sms, stream);
        SLZW_LANE_CONFIGS(X)
#undef X
        case 1: return a.p.flavour == SLZW_FLAVOUR_FIXED ? EncLat::launch(a, num_sms, stream)
                                                         : Enc0::launch(a, num_sms, stream);
        case 2: return Enc0::launch(a, num_sms, stream);
        // fixed flavour, no more streams than the latency variant has warps on the device: the
        // batch is bound by the chain of its longest stream
        default: return (a.p.flavour == SLZW_FLAVOUR_FIXED && a.n <= (uint64_t)num_sms * 12u)
                            ? EncLat::launch(a, num_sms, stream)
                            : Enc0::launch(a, num_sms, stream);
    }
}

}  // namespace slzw
