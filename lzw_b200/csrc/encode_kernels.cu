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
// x' = x * 0x9E5 mod 4096 is a bijection of the 12-bit codes ("scrambled" codes): dense sequential
// codes would form runs under an XOR hash, scrambled ones do not, and prefix' ^ hash(byte) then
// collides no more often than a 32-bit multiplicative hash of the whole key (measured,
// profiles/r01_encode_notes.md).  Numbering of new entries is insertion order and lookups are
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
#ifndef SLZW_HIT_REDUX
#define SLZW_HIT_REDUX 1  // hit detection: one warp min-reduction (1) or ballot + find-first + shuffle (0)
#endif


namespace slzw {

namespace {

constexpr int kSlots = 4096;
constexpr uint32_t kIdxMask4 = (uint32_t)(kSlots - 1) << 2;  // byte offset of a slot
constexpr uint32_t kBucketMask = 127u << 7;                   // byte offset of a 32-slot bucket
constexpr uint32_t kScr = 0x9E5u;     // code -> code' = code * kScr mod 4096 (odd => bijection)
constexpr uint32_t kScrInv = 0xBEDu;  // kScr * kScrInv == 1 mod 4096
constexpr uint32_t kByteMul = 0x6A7u; // byte -> hash contribution
static_assert(((kScr * kScrInv) & 0xFFFu) == 1u, "kScrInv must invert kScr mod 4096");

__device__ __forceinline__ uint32_t scr(uint32_t code) { return (code * kScr) & 0xFFFu; }
__device__ __forceinline__ uint32_t unscr(uint32_t code) { return (code * kScrInv) & 0xFFFu; }
// The bucket lookups (match_tile_bucket) keep codes as q = code' - 1 mod 4096 and store slots
// complemented, ~(key | q), so that an empty slot is still 0: slot ^ ~key is then below 4095
// exactly for the slot that holds the key (q == 4095 would be code 0, which is never a dictionary
// value; an empty slot gives ~key, at least 4095), so ONE warp reduction, min over the 32 slots of
// slot ^ ~key, answers "found?" and returns the new prefix at once -- no compare + ballot +
// find-first + shuffle (profiles/r02_encode_notes.md).
template <bool Q>
__device__ __forceinline__ uint32_t scrq(uint32_t code) { return (code * kScr - (Q ? 1u : 0u)) & 0xFFFu; }
template <bool Q>
__device__ __forceinline__ uint32_t unscrq(uint32_t e) { return (e * kScrInv + (Q ? kScrInv : 0u)) & 0xFFFu; }

// Dictionary accesses by 32-bit shared-window address.  They are volatile asm statements so that
// the compiler keeps them exactly where the match loop puts them (in particular the speculative
// probe stays ahead of the branch it speculates on) and in order with each other.
__device__ __forceinline__ uint32_t tbl_ld(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(v) : "r"(saddr));
    return v;
}
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
    alignas(16) uint2 rec[kRecs];  // per input byte: {byte << 12, hash bits | table base}
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

// Collision path of the dictionary lookup.  The home slot (address `a`) held another key.
// Every lane of the warp then looks at one of the next 32 slots of the linear-probing sequence;
// two ballots give the first matching and the first empty slot, and whichever comes first decides
// (an entry is never stored past an empty slot of its own probe sequence).  One shared-memory
// wavefront resolves what would be up to 32 dependent probes, which is what lets the table run at
// a load factor of 0.94 (4096 slots for <= 3838 entries).
// Returns hit; `a` = address of the matching or of the empty slot, `s` = its word.
__device__ __forceinline__ bool probe_wide(uint32_t tb, uint32_t key, int lane, uint32_t& a,
                                           uint32_t& s) {
    uint32_t h4 = a + 4u * (uint32_t)(lane + 1);  // this lane's slot in the first window
    // the table always keeps empty slots (<= 4091 entries); the bound only keeps a corrupted
    // table from hanging the warp
#pragma unroll 1
    for (int round = 0; round < kSlots / kWarpSize + 1; round++) {
        const uint32_t v = tbl_ld(tb | (h4 & kIdxMask4));
        const uint32_t bm = __ballot_sync(kFullMask, ((v ^ key) >> 12) == 0u && v != 0u);
        const uint32_t be = __ballot_sync(kFullMask, v == 0u);
        const uint32_t stop = bm | be;
        if (stop) {
            const int pos = __ffs(stop) - 1;
            a = tb | (__shfl_sync(kFullMask, h4, pos) & kIdxMask4);
            s = __shfl_sync(kFullMask, v, pos);
            return (bm >> pos) & 1u;
        }
        h4 += 4u * kWarpSize;
    }
    s = 0u;
    return false;
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

// The match loop of encoder.rs:313-337 / 639-651 over the `len` byte records of one tile.
// rec[i] = {byte << 12, table base | hash contribution of the byte as a slot byte offset}.
// The loop never leaves the tile early: input validation truncates the tile beforehand and the
// capacity check happens per tile (see encode_stream).
//
template <bool FIXED, int U>
__device__ __forceinline__ void match_tile(uint32_t* __restrict__ table, const uint32_t tb,
                                           const uint2* __restrict__ rec,
                                           uint16_t* __restrict__ codes, const int lane,
                                           const uint32_t len, MatchState& m, const uint32_t cs,
                                           const uint32_t inc, const uint32_t clear_code,
                                           const uint32_t first_code) {
    uint32_t t = m.t;
    uint32_t ncs = m.ncs;
    uint32_t ws = FIXED ? 12u : m.ws;
    uint32_t wtag = ws << 12;
    uint32_t mask = m.mask;
    uint32_t until = m.until;
    uint32_t cp = m.ncodes;
    uint32_t i = 0;
    uint2 r0;
    uint32_t a, s;

    // new index == mask (encoder.rs:326): width bump or clear + dictionary restart
#define SLZW_BUMP()                                                                             \
    {                                                                                           \
        if (ws < 12u) { /* encoder.rs:327-328 */                                                \
            ws++;                                                                               \
            wtag = ws << 12;                                                                    \
            const uint32_t nm = (1u << ws) - inc;                                               \
            until = nm - mask;                                                                  \
            mask = nm;                                                                          \
        } else { /* encoder.rs:329-333: clear at 12 bits, dictionary restarts */                \
            codes[cp++] = (uint16_t)(scr(clear_code) | (12u << 12));                            \
            ws = cs + 1u;                                                                       \
            wtag = ws << 12;                                                                    \
            mask = (1u << ws) - inc;                                                            \
            until = mask - first_code + 1u;                                                     \
            ncs = scr(first_code);                                                              \
            __syncwarp();                                                                       \
            clear_table(table, lane);                                                           \
            __syncwarp();                                                                       \
        }                                                                                       \
    }

    // One byte of encoder.rs:313-337.  RC = record of this byte, RN = record
    // of the next byte (anything addressable when this is the last byte of the tile: the
    // lookahead is discarded).  On entry `s` is the word of the home slot `a` of the key (t, RC).
#define SLZW_STEP(RC, RN)                                                                       \
    {                                                                                           \
        const uint32_t an = ((s << 2) & kIdxMask4) ^ (RN).y;                                    \
        const uint32_t sn = tbl_ld(an);    /* speculative: assumes this byte hits */            \
        const uint32_t x = s ^ t ^ (RC).x; /* == code' iff the slot holds this key */           \
        if (x - 1u < 4095u) {              /* find_word hit, encoder.rs:319-320 */              \
            t = s << 20;                                                                        \
            s = sn;                                                                             \
            a = an;                                                                             \
        } else {                                                                                \
            const uint32_t key = t | (RC).x;                                                    \
            bool hit = false;                                                                   \
            if (s != 0u) hit = probe_wide(tb, key, lane, a, s);                                 \
            if (hit) {                                                                          \
                t = s << 20;                                                                    \
            } else {                                                                            \
                /* miss: encoder.rs:322-324 / 645-649 */                                        \
                codes[cp++] = (uint16_t)((t >> 20) | wtag);                                     \
                if (!FIXED || until != 0u) {                                                    \
                    tbl_st(a, key | ncs);                                                       \
                    ncs = (ncs + kScr) & 0xFFFu;                                                \
                    until--;                                                                    \
                    if (!FIXED && until == 0u) SLZW_BUMP()                                      \
                }                                                                               \
                t = (RC).x * (kScr << 8); /* prefix = this byte: (k * kScr mod 4096) << 20 */   \
            }                                                                                   \
            a = ((t >> 18) & kIdxMask4) ^ (RN).y;                                               \
            s = tbl_ld(a);                                                                      \
        }                                                                                       \
    }

    r0 = rec[0];
    a = (t >> 18) ^ r0.y;  // home slot of (prefix, byte 0); t >> 18 == prefix' << 2
    s = tbl_ld(a);
    // U = unrolling of the step (loop overhead against instruction-cache footprint)
    if constexpr (U >= 4) {
        while (i + 4u <= len) {
            const uint2 r1 = rec[i + 1], r2 = rec[i + 2], r3 = rec[i + 3], r4 = rec[i + 4];
            SLZW_STEP(r0, r1)
            SLZW_STEP(r1, r2)
            SLZW_STEP(r2, r3)
            SLZW_STEP(r3, r4)
            r0 = r4;
            i += 4u;
        }
    } else if constexpr (U >= 2) {
        while (i + 2u <= len) {
            const uint2 r1 = rec[i + 1], r2 = rec[i + 2];
            SLZW_STEP(r0, r1)
            SLZW_STEP(r1, r2)
            r0 = r2;
            i += 2u;
        }
    }
#pragma unroll 1
    while (i < len) {
        const uint2 r1 = rec[i + 1];
        SLZW_STEP(r0, r1)
        r0 = r1;
        i += 1u;
    }
#undef SLZW_STEP
#undef SLZW_BUMP
    m.t = t;
    m.ncs = ncs;
    m.ws = ws;
    m.mask = mask;
    m.until = until;
    m.ncodes = cp;
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
// rec[i] = {byte << 12, hash7(byte) << 7} (shared memory) or {byte << 12, dictionary's
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
    uint32_t t = m.t;                             // prefix' << 20: key position
    uint32_t tp = TMEM ? m.t >> 20 : m.t >> 13;   // prefix' (column) / prefix' << 7 (bucket offset)
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
        /* slots hold ~(key | q): an empty slot is 0 and slot ^ ~key is q for the slot that   \
           holds the key, at least 4095 for every other one */                                   \
        const uint32_t key = ~(t | ((RC).x & 0xFF000u));                                        \
        uint32_t a = TMEM ? ((tp & 0x7Fu) ^ (RC).y) : (tlr | ((tp ^ (RC).y) & kBucketMask));    \
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
                const uint32_t c = __reduce_min_sync(kFullMask, v ^ key);                       \
                if (c < 4095u) { /* find_word hit, encoder.rs:319-320 */                        \
                    t = c << 20;                                                                \
                    tp = TMEM ? c : c << 7;                                                     \
                    break;                                                                      \
                }                                                                               \
            } else {                                                                            \
                const uint32_t x = v ^ key;                                                     \
                const uint32_t bal = __ballot_sync(kFullMask, x < 4095u);                       \
                if (bal) {                                                                      \
                    const uint32_t c = __shfl_sync(kFullMask, x, (int)bfind(bal));              \
                    t = c << 20;                                                                \
                    tp = TMEM ? c : c << 7;                                                     \
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
            /* the prefix' being emitted: tp itself in the tensor-memory variant */             \
            sts_u16(cpa, MODE == 1 ? ((TMEM ? tp : t >> 20) | wtag) : (TMEM ? tp : t >> 20));   \
            cpa += 2u;                                                                          \
            if (MODE == 0 || (MODE == 1 && (!FIXED || until != 0u))) {                          \
                const bool mine = em == lanes_ge; /* I own the first empty slot */              \
                const uint32_t entry = key & ~(ncs & 0xFFFu);                                   \
                if (TMEM) {                                                                     \
                    tmem_st(a, mine ? entry : v);                                               \
                    tmem_wait_st();                                                             \
                } else {                                                                        \
                    if (mine) tbl_st(a ^ tlx, entry);                                           \
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
            /* prefix = this byte: its q sits in the low bits of the record (only the low 7   \
               bits of tp / bits 7..13 of tp << 7 reach the bucket address) */                   \
            t = (RC).x << 20;                                                                   \
            tp = TMEM ? ((RC).x & 0xFFFu) : (RC).x << 7;                                        \
            break;                                                                              \
        }                                                                                       \
    }

    uint32_t i = 0;
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
    m.t = t;
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
__device__ __forceinline__ uint32_t load_tile_word(const uint8_t* __restrict__ p, uint32_t skew,
                                                   uint32_t tile_len, int lane) {
    const uint32_t lo = 4u * (uint32_t)lane;
    const uint32_t end = skew + tile_len;
    if (lo >= skew && lo + 4u <= end) return __ldg(reinterpret_cast<const uint32_t*>(p - skew + lo));
    uint32_t v = 0u;
    if (lo + 3u >= skew && lo < end) {
#pragma unroll
        for (uint32_t b = 0; b < 4u; b++)
            if (lo + b >= skew && lo + b < end) v |= (uint32_t)__ldg(p - skew + lo + b) << (8u * b);
    }
    return v;
}

// tmem (warp-uniform) = the stream's dictionary lives in tensor memory at address `tb` (table is
// unused); otherwise `table` / `tb` are the generic pointer and the shared-window address of its
// 16 KB.  One function body for both kinds of warp: everything but the match loop is shared, which
// keeps the instruction footprint of the 28 warps down.
template <int TILE, bool FIXED, bool HAS_TMEM, int U, bool BS>
__device__ void encode_stream(const DevBatch& a, uint32_t sid, uint32_t* __restrict__ table,
                              const uint32_t tb, const bool tmem_warp, EncMisc<TILE>& S, int lane) {
    const bool TMEM = HAS_TMEM && tmem_warp;
    constexpr bool Q = BS || HAS_TMEM;  // bucket lookups: codes as q = code' - 1, empty slots all ones
    using Misc = EncMisc<TILE>;
    static_assert(TILE % 4 == 0 && TILE <= 4 * kWarpSize, "one 32-bit word per lane");
    uint2* __restrict__ rec = S.rec;
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
    m.ncs = scrq<Q>(first_code);
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
        codes[m.ncodes++] = (uint16_t)(scrq<Q>(code) | (width << 12));
    };

    // Whole warp: bit-pack the buffered codes, flush complete words.
    // An entry without a width tag (stored by a match loop that cannot meet a width change) takes
    // the width `wdef`.
    auto pack_and_flush = [&](uint32_t count, uint32_t wdef) {
        for (uint32_t base = 0; base < count; base += kWarpSize) {
            const uint32_t idx = base + lane;
            const uint32_t e = idx < count ? codes[idx] : 0u;
            const uint32_t wd = idx < count ? ((e >> 12) ? (e >> 12) : wdef) : 0u;
            const uint32_t code = unscrq<Q>(e) & ((1u << wd) - 1u);  // BitWriter masks to the width
            // inclusive prefix sum of the widths; codes without a tag all have the tile's width,
            // which is the common case (MODE 0 / 2 tiles) and needs no scan
            uint32_t x, total;
            if (!__any_sync(kFullMask, (e >> 12) != 0u)) {
                const uint32_t left = count - base;
                x = ((uint32_t)lane + 1u) * wdef;
                total = (left < (uint32_t)kWarpSize ? left : (uint32_t)kWarpSize) * wdef;
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
        const uint32_t first = __ldg(src);  // encoder.rs:311 / 637: not range-checked
        m.t = scrq<Q>(first) << 20;
        if (!FIXED && n > 1 && first >= first_code) {
            // find_word would index past tree.nodes (encoder.rs:99) unless the second byte
            // is rejected first (encoder.rs:315-317)
            const uint32_t k = __ldg(src + 1);
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
                // bucket lookups: the low 12 bits carry q of the byte itself (the prefix after a miss)
                rec[idx] = make_uint2((k << 12) | (Q ? scrq<true>(k) : 0u), TMEM ? (tb | h7)
                                               : BS ? (h7 << 7)
                                                    : (tb | (((k * kByteMul) << 2) & kIdxMask4)));
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
        if (len) {
            // inserts left before the next event against the most this tile can insert
            const int mode = (FIXED && m.until == 0u) ? 2 : (m.until > len ? 0 : 1);
            // unroll of the MODE 0 and MODE 2 loops: 4 steps per iteration amortise the loop's own
            // instructions (81.9 against 85.0 ms at config 3, 51.7 against 53.7 ms at config 5);
            // 8 spill and run at 90.6 ms.  MODE 1 tiles are rare and keep U.
            constexpr int kU0 = SLZW_U0;
#define SLZW_MATCH_B(TM, MD)                                                                        \
    match_tile_bucket<FIXED, TM, (MD == 1 ? U : kU0), MD>(table, TM ? 0u : (tb | (4u * (uint32_t)lane)), TM ? tb : 0u, rec, \
                                        codes, lane, len, m, cs, inc, clear_code, first_code)
            if (TMEM) {
                if (mode == 0) SLZW_MATCH_B(true, 0);
                else if (mode == 1) SLZW_MATCH_B(true, 1);
                else if constexpr (FIXED) SLZW_MATCH_B(true, 2);
            } else if constexpr (BS) {
                if (mode == 0) SLZW_MATCH_B(false, 0);
                else if (mode == 1) SLZW_MATCH_B(false, 1);
                else if constexpr (FIXED) SLZW_MATCH_B(false, 2);
            } else {
                match_tile<FIXED, U>(table, tb, rec, codes, lane, len, m, cs, inc, clear_code, first_code);
            }
#undef SLZW_MATCH_B
        }
        __syncwarp();

        pack_and_flush(m.ncodes, ws_tile);
        m.ncodes = 0;
        // The `&mut [u8]` writer fails at the first byte past the slot (io.rs:244 / 307).  Codes
        // are emitted in input order, so a write failure inside this tile precedes a rejected
        // byte that ended it.
        if (bits > limit) {
            status = SLZW_ERR_IO_WRITE_ZERO;
        } else if (len < tile_len) {
            status = SLZW_ERR_UNEXPECTED_CODE;
            detail = (rec[len].x >> 12) & 0xFFu;
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

// ---- thread-per-stream matcher ("lanes") --------------------------------------------------------
// One LANE per stream: a warp advances up to 32 independent streams with one instruction stream,
// which is what the warp-per-stream matcher above cannot do (every one of its instructions serves
// one stream).  A lane can only own a dictionary it can address on its own, i.e. one in shared
// memory (as many lanes of ONE warp as 16 KB tables fit beside the rest) or in global memory (all
// 32 lanes of further warps; 16 KB per lane, L2-resident while their total stays below the L2).
//
// Dictionary of a lane: 512 buckets of 8 slots (32 bytes: one sector / two 128-bit loads), slot =
// [prefix':12 | byte:8 | code':12] as above.  home bucket = (prefix' ^ h9(byte)) & 511, a full
// bucket continues at bucket + step(byte) (odd step: double hashing over buckets); a bucket fills
// from slot 0 upwards, so "key not in the bucket and slot 7 empty" is a miss and the first empty
// slot takes the new entry.  1.18 buckets per input byte on the config-3 strips at the format's
// load factor of up to 0.94 (tools/exp/probe_sim.c; 4.8 slots per byte with linear probing over
// single slots).
//
// One iteration of the loop is one bucket probe of every lane, whatever the lane's outcome (hit:
// next byte; miss: insert, emit, next byte; full bucket: next bucket), so no lane waits for the
// probe sequence of another.  Codes are packed by the lane itself into a 64-bit accumulator and
// leave as aligned 32-bit stores; input bytes come from a per-lane ring in shared memory that
// cp.async fills RCH - 1 chunks of 16 bytes ahead of the lane (no registers, no stall).  Rare events
// (stream start / end, width bump, dictionary reset, errors) raise a per-lane flag and are served
// between iterations; dictionary clears are done by the whole warp.
constexpr uint32_t kLaneBucketMask = 0x3FE0u;  // byte offset of a 32-byte bucket inside a table
constexpr uint32_t kLaneHashMul = 0x6A7u << 5;  // byte -> bucket byte offset
constexpr uint32_t kLaneStepMul = 0x9Bu << 3;   // byte -> probe step (bucket byte offset, made odd)

enum : uint32_t { LS_NONE = 0, LS_FETCH, LS_EVENT, LS_END, LS_BAD_BYTE, LS_OVERFLOW, LS_SEGMENT };

template <bool GLOBAL>
__device__ __forceinline__ uint4 lane_ld(uint32_t lo, uint32_t hi) {
    uint4 v;
    if (GLOBAL) {
        asm volatile("{ .reg .b64 a; mov.b64 a, {%4, %5}; ld.global.cg.v4.u32 {%0, %1, %2, %3}, [a]; }\n"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "r"(lo), "r"(hi)
                     : "memory");
    } else {
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "r"(lo)
                     : "memory");
    }
    return v;
}
template <bool GLOBAL>
__device__ __forceinline__ void lane_st(uint32_t lo, uint32_t hi, uint32_t v) {
    if (GLOBAL) {
        asm volatile("{ .reg .b64 a; mov.b64 a, {%0, %1}; st.global.cg.u32 [a], %2; }\n" ::"r"(lo), "r"(hi), "r"(v)
                     : "memory");
    } else {
        asm volatile("st.shared.u32 [%0], %1;\n" ::"r"(lo), "r"(v) : "memory");
    }
}
template <bool GLOBAL>
__device__ __forceinline__ void lane_st_zero16(uint32_t lo, uint32_t hi) {
    const uint32_t z = 0u;
    if (GLOBAL) {
        asm volatile("{ .reg .b64 a; mov.b64 a, {%0, %1}; st.global.cg.v4.u32 [a], {%2, %2, %2, %2}; }\n" ::"r"(lo),
                     "r"(hi), "r"(z)
                     : "memory");
    } else {
        asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};\n" ::"r"(lo), "r"(z) : "memory");
    }
}

// Output side of a lane: bit accumulator + position in the slot.
struct LaneOut {
    uint64_t acc;   // LSB-first: pending bits at the bottom; MSB-first: at the top
    uint32_t nb;    // pending bits
    uint8_t* outp;  // address of the next aligned word (starts at dst - mis)
    uint32_t room;  // aligned words that still fit into the slot (saturating, re-armed when 0)
    uint32_t hold;  // the first word of a misaligned slot is kept back: its leading bytes belong
    uint32_t w0;    //   to the neighbouring slot (written byte by byte when the stream ends)
};

// BitWriter::write (io.rs:234-248 / 296-311).  Returns false when a complete word no longer fits
// into the slot: the accumulator keeps the bits and the caller ends the stream.
__device__ __forceinline__ bool lane_emit(LaneOut& o, uint32_t code, uint32_t width, const bool big,
                                          const uint32_t prmt_sel, const bool has_out) {
    const uint32_t sh = big ? 64u - o.nb - width : o.nb;
    o.acc |= (uint64_t)code << sh;
    o.nb += width;
    if (o.nb >= 32u) {
        if (o.room == 0u) return false;
        const uint32_t w = __byte_perm((uint32_t)o.acc, (uint32_t)(o.acc >> 32), prmt_sel);
        if (o.hold) {
            o.w0 = w;
            o.hold = 0u;
        } else if (has_out) {
            *reinterpret_cast<uint32_t*>(o.outp) = w;
        }
        o.outp += 4;
        o.room--;
        o.nb -= 32u;
        o.acc = big ? o.acc << 32 : o.acc >> 32;
    }
    return true;
}

// RCH = chunks of 16 bytes in a lane's input ring (a power of two, at least 2).
template <bool FIXED, bool GLOBAL, int RCH>
__device__ void encode_lanes(const DevBatch& a, const uint32_t tbl_lo, const uint32_t tbl_hi,
                             const uint32_t ring_s, const bool enabled, const int lane) {
    constexpr uint32_t kRingMask = 16u * RCH - 1u;
    constexpr uint32_t F_FETCH = 1u;  // the lane wants a stream
    constexpr uint32_t F_EVENT = 2u;  // new index == mask (encoder.rs:326)
    constexpr uint32_t F_STOP = 4u;   // the match loop cannot go on: input (segment) consumed, a
                                      // word does not fit into the slot, or a rejected byte
    const bool big = a.p.big_endian != 0;
    const uint32_t prmt_sel = big ? 0x4567u : 0x3210u;
    const uint32_t inc = (!FIXED && a.p.tiff_early_change) ? 1u : 0u;
    const bool has_out = a.out != nullptr;

    uint32_t svc = enabled ? F_FETCH : 0u;
    bool act = false;       // a stream is in the match loop
    bool need_adv = false;  // the current byte is consumed, the next lookup is not set up yet
    // stream
    uint32_t sid = 0, cs = 8, max_code = 255, status = SLZW_OK, detail = 0;
    const uint8_t* in_end = nullptr;  // one past the last input byte
    const uint8_t* seg_ptr = nullptr; // address of the current byte when `left` was last armed
    uint32_t seg_len = 0;             // value `left` was armed with
    uint8_t* dst = nullptr;
    uint64_t cap = 0;
    uint32_t mis = 0;
    // matcher
    uint32_t p = 0;                 // current_prefix' (encoder.rs:311)
    uint32_t key = 0, a0 = tbl_lo;  // key and bucket of the pending lookup
    uint32_t stp = 32, pb = 0;      // probe step and root code' of the current byte
    uint32_t bn = 0;                // the byte after the current one
    uint32_t left = 0;              // input bytes after the current one (this segment)
    uint32_t g32 = 0;               // low address bits of the current byte
    const uint8_t* pf = nullptr;    // next chunk the ring fetches
    uint32_t ncs = 0, until = 0, ws = 12, wmask = 0xFFF, mask = 0;
    LaneOut o;
    o.acc = 0; o.nb = 0; o.outp = nullptr; o.room = 0; o.hold = 0; o.w0 = 0;

    auto ring_fetch = [&]() {  // one chunk into its ring slot (nothing past the stream's last chunk)
        if (pf < in_end)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(ring_s + ((uint32_t)(uintptr_t)pf & kRingMask)),
                         "l"(pf)
                         : "memory");
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        pf += 16;
    };
    auto ring_byte = [&](uint32_t addr_lo) -> uint32_t {
        uint32_t v;
        asm volatile("ld.shared.u8 %0, [%1];\n" : "=r"(v) : "r"(ring_s + (addr_lo & kRingMask)) : "memory");
        return v;
    };
    auto words_left = [&]() -> uint64_t {  // aligned words that fit behind outp
        if (!has_out) return ~0ull;
        const uint64_t total = ((uint64_t)mis + cap) >> 2;
        const uint64_t used = (uint64_t)(o.outp - (dst - mis)) >> 2;
        return total > used ? total - used : 0ull;
    };
    // a word that lane_emit could not flush: re-arm the 32-bit word budget if the slot has room
    auto flush_pending = [&]() -> bool {
        while (o.nb >= 32u) {
            const uint64_t wl = words_left();
            if (wl == 0ull) return false;
            o.room = (uint32_t)(wl < 0xFFFFFFFFull ? wl : 0xFFFFFFFFull);
            const uint32_t w = __byte_perm((uint32_t)o.acc, (uint32_t)(o.acc >> 32), prmt_sel);
            if (o.hold) {
                o.w0 = w;
                o.hold = 0u;
            } else if (has_out) {
                *reinterpret_cast<uint32_t*>(o.outp) = w;
            }
            o.outp += 4;
            o.room--;
            o.nb -= 32u;
            o.acc = big ? o.acc << 32 : o.acc >> 32;
        }
        return true;
    };

    for (;;) {
        // ---- service: stream ends / starts, width bumps, dictionary resets, errors ----
        if (__any_sync(kFullMask, svc != 0u)) {
            bool clear = false;
            if (svc & (F_EVENT | F_STOP)) {
                bool alive = flush_pending();  // false: the bit writer failed (io.rs:244 / 307)
                if (alive && (svc & F_EVENT)) {
                    if (ws < 12u) {  // encoder.rs:327-328
                        ws++;
                        const uint32_t nm = (1u << ws) - inc;
                        until = nm - mask;
                        mask = nm;
                        wmask = (1u << ws) - 1u;
                    } else {         // encoder.rs:329-333: clear code at 12 bits, dictionary restarts
                        const uint32_t first_code = (1u << cs) + 2u;
                        lane_emit(o, 1u << cs, 12u, big, prmt_sel, has_out);
                        alive = flush_pending();
                        ws = cs + 1u;
                        wmask = (1u << ws) - 1u;
                        mask = (1u << ws) - inc;
                        until = mask - first_code + 1u;
                        ncs = scr(first_code);
                        clear = true;
                    }
                }
                svc &= ~F_EVENT;
                if (!alive) {
                    if (status == SLZW_OK) status = SLZW_ERR_IO_WRITE_ZERO;
                    svc = F_FETCH;
                } else if (status != SLZW_OK) {
                    svc = F_FETCH;  // rejected byte
                } else if (svc & F_STOP) {
                    // input consumed: the whole stream, or a 32-bit segment of it
                    const uint8_t* c = seg_ptr + seg_len;  // current byte (left == 0)
                    const uint64_t rest = need_adv && left == 0u ? (uint64_t)(in_end - c) - 1u : 1u;
                    if (rest != 0u) {
                        if (need_adv && left == 0u) {
                            seg_len = left = (uint32_t)(rest < 0x40000000ull ? rest : 0x40000000ull);
                            seg_ptr = c;
                        }
                        svc = 0u;  // back to the match loop
                    } else {
                        svc = F_FETCH;
                    }
                }
            }
            if (svc & F_FETCH) {
                // finish the current stream (if any), then take streams from the queue until one
                // needs the match loop or the queue is empty
                bool have = act;
                bool has_prefix = act;
                act = false;
                for (;;) {
                    if (have) {
                        bool finished = false;
                        if (status == SLZW_OK) {
                            // encoder.rs:339-343 / 653-655 (300-309 for an empty stream)
                            finished = true;
                            if (has_prefix) {
                                lane_emit(o, unscr(p) & wmask, ws, big, prmt_sel, has_out);
                                finished = flush_pending();
                            }
                            if (finished && !FIXED) {
                                lane_emit(o, ((1u << cs) + 1u) & wmask, ws, big, prmt_sel, has_out);
                                finished = flush_pending();
                            }
                            if (!finished) status = SLZW_ERR_IO_WRITE_ZERO;
                        }
                        // whole bytes the bit writer has seen; it fails when it reaches byte index `cap`
                        const int64_t flushed = (int64_t)(o.outp - dst);
                        if (status != SLZW_OK && status != SLZW_ERR_CODE_SIZE &&
                            (uint64_t)(flushed + (int64_t)(o.nb >> 3)) > cap) {
                            status = SLZW_ERR_IO_WRITE_ZERO;  // the writer failed before the byte was read
                            detail = 0;
                        }
                        uint64_t total = (uint64_t)(flushed + (int64_t)((finished ? o.nb + 7u : o.nb) >> 3));
                        if (finished && total > cap) status = SLZW_ERR_IO_WRITE_ZERO;  // fill(), io.rs:251-259
                        if (total > cap) total = cap;
                        if (has_out) {
                            const uint32_t nbytes = (o.nb + 7u) >> 3;
                            for (uint32_t j = 0; j < nbytes; j++) {
                                const int64_t at = flushed + (int64_t)j;
                                const uint32_t v = big ? (uint32_t)(o.acc >> (56u - 8u * j)) : (uint32_t)(o.acc >> (8u * j));
                                if (at >= 0 && (uint64_t)at < total) dst[at] = (uint8_t)v;
                            }
                            if (mis != 0u && o.hold == 0u)
                                for (uint32_t j = mis; j < 4u; j++)
                                    if ((uint64_t)(j - mis) < total) dst[j - mis] = (uint8_t)(o.w0 >> (8u * j));
                        }
                        a.out_len[sid] = total;
                        a.status[sid] = status;
                        a.detail[sid] = detail;
                        atomicAdd(a.queue + (GLOBAL ? 7 : 6), a.in_off[sid + 1] - a.in_off[sid]);
                        have = false;
                    }
                    // ---- next stream ----
                    const unsigned long long q = atomicAdd(a.queue, 1ull);
                    if (q >= a.n) {
                        svc = 0u;
                        break;
                    }
                    have = true;
                    has_prefix = false;
                    sid = a.order ? a.order[q] : (uint32_t)q;
                    const uint64_t in_begin = a.in_off[sid];
                    const uint64_t n = a.in_off[sid + 1] - in_begin;
                    const uint8_t* src = a.in + in_begin;
                    in_end = src + n;
                    dst = nullptr;
                    cap = ~0ull;
                    if (has_out) {
                        const uint64_t ob = a.out_off[sid];
                        dst = a.out + ob;
                        cap = a.out_off[sid + 1] - ob;
                    }
                    mis = (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 3u);
                    cs = FIXED ? 8u : (a.code_size ? a.code_size[sid] : a.p.code_size);
                    status = SLZW_OK;
                    detail = 0;
                    o.acc = 0;
                    o.nb = 8u * mis;
                    o.outp = dst - mis;
                    o.hold = mis != 0u ? 1u : 0u;
                    o.room = 0u;  // armed by the first flush_pending
                    if (!FIXED && (cs < 2u || cs > 8u)) {  // encoder.rs:281-283: nothing is written
                        status = SLZW_ERR_CODE_SIZE;
                        detail = cs;
                        o.nb = 0;
                        o.outp = dst;
                        o.hold = 0;
                        mis = 0;
                        continue;
                    }
                    max_code = (1u << cs) - 1u;  // encoder.rs:285
                    const uint32_t first_code = FIXED ? 256u : (1u << cs) + 2u;
                    ws = FIXED ? 12u : cs + 1u;  // encoder.rs:289
                    wmask = (1u << ws) - 1u;
                    mask = (1u << ws) - inc;     // encoder.rs:292
                    until = FIXED ? 4096u - first_code : mask - first_code + 1u;
                    ncs = scr(first_code);
                    if (!FIXED) {                // encoder.rs:297
                        lane_emit(o, 1u << cs, ws, big, prmt_sel, has_out);
                        if (!flush_pending()) {
                            status = SLZW_ERR_IO_WRITE_ZERO;
                            continue;
                        }
                    }
                    if (n == 0) continue;  // encoder.rs:300-309: end code, fill
                    const uint32_t first = __ldg(src);  // encoder.rs:311 / 637: not range-checked
                    p = scr(first);
                    has_prefix = true;
                    if (!FIXED && n > 1 && first >= first_code) {
                        // find_word would index past tree.nodes (encoder.rs:99) unless the second
                        // byte is rejected first (encoder.rs:315-317)
                        const uint32_t k = __ldg(src + 1);
                        if (k > max_code) {
                            status = SLZW_ERR_UNEXPECTED_CODE;
                            detail = k;
                        } else {
                            status = SLZW_ERR_REFERENCE_PANIC;
                        }
                        continue;
                    }
                    if (n == 1) continue;
                    // the stream enters the match loop in front of byte 1
                    const uint64_t rest = n - 1;
                    seg_len = left = (uint32_t)(rest < 0x40000000ull ? rest : 0x40000000ull);
                    seg_ptr = src - seg_len + seg_len;  // == src: position = seg_ptr + (seg_len - left)
                    seg_ptr = src;
                    g32 = (uint32_t)reinterpret_cast<uintptr_t>(src);
                    pf = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(src + 1) & ~uintptr_t(15));
                    for (int c = 0; c < RCH; c++) ring_fetch();
                    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
                    bn = ring_byte(g32 + 1u);
                    // (g32 & 15) == 15: byte 1 starts a chunk and the ring already holds RCH chunks from there
                    need_adv = true;
                    act = true;
                    clear = true;
                    svc = 0u;
                    break;
                }
            }
            // dictionary clears, by the whole warp, one lane's table after the other
            uint32_t cm = __ballot_sync(kFullMask, clear);
            while (cm) {
                const int l = __ffs(cm) - 1;
                cm &= cm - 1u;
                const uint32_t lo = __shfl_sync(kFullMask, tbl_lo, l);
                const uint32_t hi = __shfl_sync(kFullMask, tbl_hi, l);
#pragma unroll 4
                for (uint32_t j = 0; j < kSlots * 4u / 16u / kWarpSize; j++)
                    lane_st_zero16<GLOBAL>(lo + 16u * (j * kWarpSize + (uint32_t)lane), hi);
            }
            __syncwarp();
            if (!__any_sync(kFullMask, act)) break;
        }

        // ---- one bucket probe per lane, branch-free ----
        // A lone warp issues in order, so whatever sits behind a divergent branch is serialised
        // with everything else; written as selects and predicated memory operations the three
        // outcomes of a probe (hit / miss / full bucket) and the step to the next byte overlap
        // (first version with branches: 920 cycles per iteration, see profiles/r02_lanes_notes.md).
        const bool go = act && svc == 0u;
        // step to the next byte (encoder.rs:313-318) for the lanes whose lookup is resolved
        const bool adv_try = go && need_adv;
        const bool adv = adv_try && left != 0u;
        if (adv_try && left == 0u) svc = F_STOP;  // input (segment) consumed
        const uint32_t b = bn;
        const uint32_t nkb = b << 12;
        const uint32_t nhb = ((b * kLaneHashMul) & kLaneBucketMask) | tbl_lo;
        const uint32_t nstp = ((b * kLaneStepMul) & kLaneBucketMask) | 32u;
        const uint32_t npb = scr(b);
        left -= adv ? 1u : 0u;
        g32 += adv ? 1u : 0u;
        key = adv ? ((p << 20) | nkb) : key;
        a0 = adv ? (((p << 5) & kLaneBucketMask) ^ nhb) : a0;
        stp = adv ? nstp : stp;
        pb = adv ? npb : pb;
        need_adv = need_adv && !adv;
        const bool edge = adv && (g32 & 15u) == 15u;  // the byte after this one starts a chunk
        const bool fetch = edge && pf < in_end;
        {
            const uint32_t ra = ring_s + ((g32 + 1u) & kRingMask);
            const uint32_t fa = ring_s + ((uint32_t)(uintptr_t)pf & kRingMask);
            asm volatile(
                "{\n"
                ".reg .pred pe, pa, pf;\n"
                "setp.ne.u32 pe, %1, 0;\n"
                "setp.ne.u32 pa, %2, 0;\n"
                "setp.ne.u32 pf, %3, 0;\n"
                "@pe cp.async.wait_group %7;\n"
                "@pa ld.shared.u8 %0, [%4];\n"
                "@pf cp.async.ca.shared.global [%5], [%6], 16;\n"
                "@pe cp.async.commit_group;\n"
                "}\n"
                : "+r"(bn)
                : "r"((uint32_t)edge), "r"((uint32_t)adv), "r"((uint32_t)fetch), "r"(ra), "r"(fa), "l"(pf),
                  "n"(RCH - 2)
                : "memory");
        }
        pf += edge ? 16 : 0;
        const bool bad = adv && b > max_code;  // encoder.rs:315-317
        status = bad ? (uint32_t)SLZW_ERR_UNEXPECTED_CODE : status;
        detail = bad ? b : detail;
        if (bad) svc = F_STOP;

        const bool probe = act && svc == 0u;
        const uint4 e0 = lane_ld<GLOBAL>(a0, tbl_hi);
        const uint4 e1 = lane_ld<GLOBAL>(a0 + 16u, tbl_hi);
        // slot ^ key == code' iff the slot holds the key (code' is never 0): the minimum of
        // (slot ^ key) - 1 over the bucket is below 4095 iff one slot does
        const uint32_t m0 = min(min((e0.x ^ key) - 1u, (e0.y ^ key) - 1u), (e0.z ^ key) - 1u);
        const uint32_t m1 = min(min((e0.w ^ key) - 1u, (e1.x ^ key) - 1u), (e1.y ^ key) - 1u);
        const uint32_t m2 = min((e1.z ^ key) - 1u, (e1.w ^ key) - 1u);
        const uint32_t m = min(min(m0, m1), m2);
        const bool found = m < 4095u;
        const bool full = e1.w != 0u;
        const bool hit = probe && found;             // find_word, encoder.rs:319-320
        const bool miss = probe && !found && !full;  // encoder.rs:322-324 / 645-649
        const bool coll = probe && !found && full;   // the key may have gone to the next bucket
        const uint32_t a_next = (a0 & ~kLaneBucketMask) | ((a0 + stp) & kLaneBucketMask);
        // miss: the first empty slot of the bucket (slots fill from 0 upwards) takes the entry
        const bool ins = miss && (!FIXED || until != 0u);
        {
            const bool up = e0.w != 0u;
            const uint32_t s0 = up ? e1.x : e0.x, s1 = up ? e1.y : e0.y, s2 = up ? e1.z : e0.z;
            const bool t = s1 == 0u;
            const uint32_t lo = t ? s0 : s2;
            const uint32_t idx = (up ? 4u : 0u) + (t ? 0u : 2u) + (lo != 0u ? 1u : 0u);
            const uint32_t sa = a0 + 4u * idx;
            if (GLOBAL) {
                asm volatile("{ .reg .pred q; .reg .b64 a; setp.ne.u32 q, %3, 0; mov.b64 a, {%0, %1}; @q st.global.cg.u32 [a], %2; }\n" ::"r"(sa),
                             "r"(tbl_hi), "r"(key | ncs), "r"((uint32_t)ins)
                             : "memory");
            } else {
                asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q st.shared.u32 [%0], %1; }\n" ::"r"(sa), "r"(key | ncs),
                             "r"((uint32_t)ins)
                             : "memory");
            }
        }
        ncs = ins ? ((ncs + kScr) & 0xFFFu) : ncs;
        until -= ins ? 1u : 0u;
        if (!FIXED && ins && until == 0u) svc |= F_EVENT;
        // miss: BitWriter::write of the prefix (io.rs:234-248 / 296-311)
        {
            const uint32_t code = unscr(p) & wmask;
            const uint32_t sh = big ? 64u - o.nb - ws : o.nb;
            const uint64_t accn = o.acc | ((uint64_t)code << (sh & 63u));
            o.acc = miss ? accn : o.acc;
            o.nb += miss ? ws : 0u;
            const bool fl = miss && o.nb >= 32u;
            const bool st = fl && o.room != 0u;  // room == 0: the service block decides
            if (fl && o.room == 0u) svc |= F_STOP;
            const uint32_t w = __byte_perm((uint32_t)o.acc, (uint32_t)(o.acc >> 32), prmt_sel);
            asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q st.global.u32 [%0], %1; }\n" ::"l"(o.outp), "r"(w),
                         "r"((uint32_t)(st && has_out))
                         : "memory");
            o.outp += st ? 4 : 0;
            o.room -= st ? 1u : 0u;
            o.nb -= st ? 32u : 0u;
            const uint64_t accs = big ? o.acc << 32 : o.acc >> 32;
            o.acc = st ? accs : o.acc;
        }
        p = hit ? m + 1u : (miss ? pb : p);
        a0 = coll ? a_next : a0;
        need_adv = need_adv || hit || miss;
    }
}

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
template <int TILE, int SWARPS, int TWARPS, int U, bool BS, bool FIXED>
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
    for (;;) {
        unsigned long long q = 0;
        if (lane == 0) q = atomicAdd(a.queue, 1ull);
        q = __shfl_sync(kFullMask, q, 0);
        if (q >= a.n) break;
        const uint32_t sid = a.order ? a.order[q] : (uint32_t)q;
        encode_stream<TILE, FIXED, (TWARPS > 0), U, BS>(a, sid, table, tb, tmem_warp, S, lane);
    }

    if constexpr (TWARPS > 0) {
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
        if (warp == 0)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512u));
    }
}


// ---- the kernel with lanes ------------------------------------------------------------------------
// Warps [0, TWARPS): one stream per warp, dictionary in tensor memory (encode_stream above).
// Warps [TWARPS, TWARPS + SWARPS): one stream per warp, dictionary in shared memory (bucket lookups).
// Then, if NSL > 0, one warp of NSL lanes with dictionaries in shared memory, then GW warps of 32
// lanes each with dictionaries in global memory (a.enc_tables, 16 KB per lane).  All take streams
// from the same queue.
//
// Shared memory: the SWARPS + NSL tables start at the first 16 KB boundary (so that
// `table | offset` needs no add); the per-warp blocks of the warp-per-stream warps and the input
// rings of the lanes fill the space in front of the first table and behind the last one.
template <int TILE, int TWARPS, int SWARPS, int NSL, int GW>
struct LaneLayout {
    static constexpr uint32_t kTable = kSlots * 4;
    static constexpr uint32_t kMisc = (uint32_t)((sizeof(EncMisc<TILE>) + 31) & ~size_t(31));
    static constexpr uint32_t kRingS = (64u * NSL + 63u) & ~63u;  // 4 chunks per lane
    static constexpr uint32_t kRingG = 32u * 32u;                  // 2 chunks per lane
    static constexpr uint32_t kHead = 64;                          // tensor-memory base address slot
    static constexpr int kWps = TWARPS + SWARPS;                   // warp-per-stream warps
    static constexpr int kLaneWarp = NSL > 0 ? 1 : 0;
    static constexpr int kItems = kWps + kLaneWarp + GW;
    static constexpr int kWarps = kItems;
    __host__ __device__ static uint32_t item_bytes(int i) {
        return i < kWps ? kMisc : (i < kWps + kLaneWarp ? kRingS : kRingG);
    }
    __host__ __device__ static uint32_t first_table(uint32_t base) { return (base + kTable - 1) & ~(kTable - 1); }
    // shared-window address of item i; `end` receives the end of the last item
    __host__ __device__ static uint32_t place(uint32_t base, int want, uint32_t* end) {
        uint32_t front = (base + 63u) & ~63u;
        const uint32_t front_hi = first_table(base);
        uint32_t back = front_hi + (SWARPS + NSL) * kTable;
        uint32_t at = 0;
        for (int i = 0; i < kItems; i++) {
            const uint32_t b = item_bytes(i);
            uint32_t where;
            if (front + b <= front_hi) {
                where = front;
                front += b;
            } else {
                where = back;
                back += b;
            }
            if (i == want) at = where;
        }
        if (end) *end = back;
        return at;
    }
    __host__ __device__ static uint32_t bytes(uint32_t base) {
        uint32_t end;
        place(base, 0, &end);
        return end - base;
    }
};

template <int TILE, int TWARPS, int SWARPS, int NSL, int GW, bool FIXED>
__global__ void __launch_bounds__(LaneLayout<TILE, TWARPS, SWARPS, NSL, GW>::kWarps * kWarpSize, 1)
slzw_encode_lanes_kernel(const DevBatch a, const uint32_t dyn_bytes) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    using L = LaneLayout<TILE, TWARPS, SWARPS, NSL, GW>;
    static_assert(TWARPS == 0 || TWARPS == 16, "tensor memory holds 16 dictionaries of 128 columns");
    const uint32_t warp = threadIdx.x / kWarpSize;
    const int lane = threadIdx.x % kWarpSize;
    const uint32_t raw = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t base = raw + L::kHead;
    if (L::bytes(base) + L::kHead > dyn_bytes) __trap();  // launch configuration and layout disagree

    uint32_t tmem_base = 0;
    if constexpr (TWARPS > 0) {
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(raw), "r"(512u));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_raw);
    }

    if (warp < (uint32_t)L::kWps) {
        if constexpr (L::kWps > 0) {
            EncMisc<TILE>& S = *reinterpret_cast<EncMisc<TILE>*>(smem_raw + (L::place(base, (int)warp, nullptr) - raw));
            const bool tmem_warp = warp < (uint32_t)TWARPS;
            uint32_t tb;
            uint32_t* table = nullptr;
            if (tmem_warp) {
                tb = tmem_base + (((warp & 3u) * 32u) << 16) + (warp >> 2) * 128u;
            } else {
                tb = L::first_table(base) + (warp - TWARPS) * L::kTable;
                table = reinterpret_cast<uint32_t*>(smem_raw + (tb - raw));
            }
            for (;;) {
                unsigned long long q = 0;
                if (lane == 0) q = atomicAdd(a.queue, 1ull);
                q = __shfl_sync(kFullMask, q, 0);
                if (q >= a.n) break;
                const uint32_t sid = a.order ? a.order[q] : (uint32_t)q;
                encode_stream<TILE, FIXED, (TWARPS > 0), 2, true>(a, sid, table, tb, tmem_warp, S, lane);
                if (lane == 0) atomicAdd(a.queue + (tmem_warp ? 4 : 5), a.in_off[sid + 1] - a.in_off[sid]);
            }
        }
    } else if (NSL > 0 && warp == (uint32_t)L::kWps) {
        const bool enabled = lane < NSL;
        const uint32_t tbl = L::first_table(base) + (SWARPS + (enabled ? (uint32_t)lane : 0u)) * L::kTable;
        const uint32_t ring = L::place(base, L::kWps, nullptr) + 64u * (uint32_t)lane;
        encode_lanes<FIXED, false, 4>(a, tbl, 0u, ring, enabled, lane);
    } else {
        if constexpr (GW > 0) {
            const uint32_t gw = warp - L::kWps - L::kLaneWarp;
            const uint64_t tbl = reinterpret_cast<uint64_t>(a.enc_tables) +
                                 (((uint64_t)blockIdx.x * GW + gw) * kWarpSize + (uint64_t)lane) * L::kTable;
            const uint32_t ring = L::place(base, L::kWps + L::kLaneWarp + (int)gw, nullptr) + 32u * (uint32_t)lane;
            encode_lanes<FIXED, true, 2>(a, (uint32_t)tbl, (uint32_t)(tbl >> 32), ring, true, lane);
        }
    }

    if constexpr (TWARPS > 0) {
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
        if (warp == 0)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512u));
    }
}

}  // namespace

// ---- launch configuration ---------------------------------------------------------------------
// {input tile, warps with a shared-memory dictionary, warps with a tensor-memory dictionary}
template <int TILE, int SWARPS, int TWARPS, int U, bool BS>
struct EncConfig {
    using L = EncLayout<TILE, SWARPS, SWARPS + TWARPS>;
    // the shared window of a CTA starts with 1 KB reserved by the system; taking the larger of
    // the two layouts keeps the launch valid should the dynamic region start at 0 instead
    static uint32_t smem() {
        const uint32_t a = L::bytes(1024u + L::kHead), b = L::bytes(L::kHead);
        return (a > b ? a : b) + L::kHead;
    }
    static cudaError_t configure() {
        cudaError_t e = cudaFuncSetAttribute(slzw_encode_kernel<TILE, SWARPS, TWARPS, U, BS, false>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem());
        if (e != cudaSuccess) return e;
        return cudaFuncSetAttribute(slzw_encode_kernel<TILE, SWARPS, TWARPS, U, BS, true>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem());
    }
    static cudaError_t launch(const DevBatch& a, int num_sms, cudaStream_t stream) {
        constexpr int WARPS = SWARPS + TWARPS;
        const uint64_t ctas = (a.n + WARPS - 1) / WARPS;
        const int grid = (int)(ctas < (uint64_t)num_sms ? ctas : (uint64_t)num_sms);
        if (a.p.flavour == SLZW_FLAVOUR_FIXED)
            slzw_encode_kernel<TILE, SWARPS, TWARPS, U, BS, true><<<grid, WARPS * kWarpSize, smem(), stream>>>(a, smem());
        else
            slzw_encode_kernel<TILE, SWARPS, TWARPS, U, BS, false><<<grid, WARPS * kWarpSize, smem(), stream>>>(a, smem());
        return cudaGetLastError();
    }
};

template <int TILE, int TWARPS, int SWARPS, int NSL, int GW>
struct LaneConfig {
    using L = LaneLayout<TILE, TWARPS, SWARPS, NSL, GW>;
    static constexpr int kWarps = L::kWarps;
    static constexpr int kStreams = TWARPS + SWARPS + NSL + 32 * GW;  // per SM
    static uint32_t smem() {
        const uint32_t a = L::bytes(1024u + L::kHead), b = L::bytes(L::kHead);
        return (a > b ? a : b) + L::kHead;
    }
    static cudaError_t configure() {
        cudaError_t e = cudaFuncSetAttribute(slzw_encode_lanes_kernel<TILE, TWARPS, SWARPS, NSL, GW, false>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem());
        if (e != cudaSuccess) return e;
        return cudaFuncSetAttribute(slzw_encode_lanes_kernel<TILE, TWARPS, SWARPS, NSL, GW, true>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem());
    }
    static size_t table_bytes(int num_sms) { return (size_t)num_sms * GW * 32u * (kSlots * 4u); }
    static cudaError_t launch(const DevBatch& a, int num_sms, cudaStream_t stream) {
        const uint64_t ctas = (a.n + kStreams - 1) / kStreams;
        const int grid = (int)(ctas < (uint64_t)num_sms ? ctas : (uint64_t)num_sms);
        if (a.p.flavour == SLZW_FLAVOUR_FIXED)
            slzw_encode_lanes_kernel<TILE, TWARPS, SWARPS, NSL, GW, true><<<grid, kWarps * kWarpSize, smem(), stream>>>(a, smem());
        else
            slzw_encode_lanes_kernel<TILE, TWARPS, SWARPS, NSL, GW, false><<<grid, kWarps * kWarpSize, smem(), stream>>>(a, smem());
        return cudaGetLastError();
    }
};

// {input tile, shared-memory dictionaries, tensor-memory dictionaries, step unrolling,
//  bucket lookups in shared memory (tensor memory always uses them)}
using Enc0 = EncConfig<96, 12, 16, 2, true>;    // 28 streams per SM, one warp each, bucket lookups
using Enc1 = EncConfig<128, 12, 0, 4, false>;   // shared memory only, scalar speculative probes
// {tile of the warp-per-stream warps, tensor-memory warps, shared-memory warps, lanes in shared
//  memory, warps of lanes in global memory}
using Lan10 = LaneConfig<64, 0, 0, 13, 0>;
using Lan11 = LaneConfig<64, 0, 0, 13, 1>;
using Lan12 = LaneConfig<64, 0, 0, 0, 1>;
using Lan13 = LaneConfig<64, 0, 0, 0, 2>;
using Lan14 = LaneConfig<64, 0, 0, 0, 4>;
using Lan15 = LaneConfig<64, 16, 0, 13, 0>;
using Lan16 = LaneConfig<80, 16, 12, 0, 1>;
using Lan17 = LaneConfig<80, 16, 12, 0, 2>;
using Lan18 = LaneConfig<80, 16, 12, 0, 4>;
using Lan19 = LaneConfig<80, 16, 12, 0, 0>;

static int g_enc_config = -1;

void encode_select_config(int c) { g_enc_config = c; }

#define SLZW_LANE_CONFIGS(X) X(10, Lan10) X(11, Lan11) X(12, Lan12) X(13, Lan13) X(14, Lan14) X(15, Lan15) X(16, Lan16) X(17, Lan17) X(18, Lan18) X(19, Lan19)

cudaError_t encode_configure() {
    cudaError_t e = Enc0::configure();
    if (e != cudaSuccess) return e;
    if ((e = Enc1::configure()) != cudaSuccess) return e;
#define X(id, C) if ((e = C::configure()) != cudaSuccess) return e;
    SLZW_LANE_CONFIGS(X)
#undef X
    return cudaSuccess;
}

static int encode_pick(uint64_t n, int num_sms) {
    int cfg = g_enc_config;
    if (cfg < 0) cfg = n <= (uint64_t)num_sms * 12u ? 1 : 0;
    return cfg;
}

// global-memory dictionaries the launch needs (0 for the configurations without such lanes)
size_t encode_table_bytes(uint64_t n, int num_sms) {
    switch (encode_pick(n, num_sms)) {
#define X(id, C) case id: return C::table_bytes(num_sms);
        SLZW_LANE_CONFIGS(X)
#undef X
        default: return 0;
    }
}

cudaError_t encode_launch(const DevBatch& a, int num_sms, cudaStream_t stream) {
    switch (encode_pick(a.n, num_sms)) {
        case 1: return Enc1::launch(a, num_sms, stream);
#define X(id, C) case id: return C::launch(a, num_sms, stream);
        SLZW_LANE_CONFIGS(X)
#undef X
        default: return Enc0::launch(a, num_sms, stream);
    }
}

}  // namespace slzw
