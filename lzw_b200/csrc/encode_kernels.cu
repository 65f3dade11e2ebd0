// encode_kernels.cu -- batched LZW encoder for sm_100a.
//
// One warp per stream, persistent CTAs (one per SM), streams handed out through a global work
// queue in the order chosen by the scheduler.  Per stream:
//   * the reference's arena trie (encoder.rs:58-149) is replaced by an open-addressing hash
//     dictionary keyed on (prefix code, next byte) -> code, one u32 per slot
//     [code:12 | prefix:12 | byte:8], resident in shared memory; numbering of new entries is
//     insertion order and lookups are exact (full linear probing), so the emitted codes equal
//     the reference's;
//   * lane 0 walks the input (the match loop is a dependent chain, one probe per byte,
//     encoder.rs:313-337) over tiles the warp stages into shared memory with cp.async, double
//     buffered; the loop is software-pipelined and predicated (see match_tile);
//   * emitted codes are buffered as [width:4 | code:12] and bit-packed by the whole warp
//     (LSB-first like io.rs:234-248 or MSB-first like io.rs:296-311) into a shared-memory word
//     window that is written to the output slot with aligned 32-bit stores;
//   * the dictionary reset (encoder.rs:329-333) is a cooperative vectorised clear.
// Semantics follow VariableEncoder::inner_encode (encoder.rs:273-346) and
// FixedEncoder::inner_encode (encoder.rs:618-658) exactly, including the unchecked first byte
// (encoder.rs:311) and the `&mut [u8]`-writer behaviour when the slot is too small.
#include "slzw_device.cuh"

namespace slzw {

namespace {

// Dictionary slot = [code:12 | prefix code:12 | byte:8]; 0 = empty (codes start at >= 6).
// The low 20 bits are the key (prefix code << 8 | byte); the slot index is a multiplicative
// hash of the key.
constexpr uint32_t kHashA = 0x9E3779B1u;

template <int SLOTS>
__device__ __forceinline__ uint32_t slot_of_key(uint32_t key) {  // key = prefix code << 8 | byte
    const uint32_t x = key * kHashA;
    if constexpr ((SLOTS & (SLOTS - 1)) == 0) {
        return x >> (32 - __builtin_ctz(SLOTS));
    } else {
        return __umulhi(x, (uint32_t)SLOTS);
    }
}

// Shared-memory accesses by 32-bit shared address (no generic-address conversion in the loop).
__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];\n" : "=r"(v) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t saddr) {
    uint32_t v;
    asm("ld.shared.u8 %0, [%1];\n" : "=r"(v) : "r"(saddr));  // input tile: read-only here
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t saddr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;\n" ::"r"(saddr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_u16(uint32_t saddr, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;\n" ::"r"(saddr), "h"((uint16_t)v) : "memory");
}

enum Reason : uint32_t { R_TILE_END = 0, R_RESET = 1, R_STOP = 2 };

template <int SLOTS, int TILE>
struct EncWarpSmem {
    static constexpr int kCodeBuf = TILE + 16;                  // codes one tile can emit
    static constexpr int kOutWords = (kCodeBuf * 12) / 32 + 4;  // packed window
    uint32_t table[SLOTS];
    uint32_t outw[kOutWords];
    uint16_t codes[kCodeBuf];
    __align__(16) uint8_t tile[2][TILE + 32];  // double-buffered input tiles
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

// Lane-0 match state (encoder.rs:289-311), kept in registers across tiles.
struct MatchState {
    uint32_t pw;          // current_prefix << 20
    uint32_t next_code;   // tree.len()
    uint32_t write_size;  // encoder.rs:289
    uint32_t mask;        // size_increase_mask, encoder.rs:292
    uint32_t ncodes;      // codes buffered for the packer
    uint32_t status, detail;
};

// Collision path of the dictionary lookup.  The first probe (slot `a`, word `s`) was neither
// the key nor empty.  Every lane of the warp then looks at one of the next 32 slots of the
// linear-probing sequence; two ballots give the first matching and the first empty slot, and
// whichever comes first decides (an entry is never stored past an empty slot of its own probe
// sequence).  One shared-memory wavefront resolves what would be up to 32 dependent probes, which
// is what lets the table run at a load factor of 0.94 (4096 slots for <= 3838 entries).
// Returns hit; `a` = shared address of the matching or of the empty slot, `s` = its word.
template <int SLOTS>
__device__ __forceinline__ bool probe_wide(uint32_t tb, uint32_t key, int lane, uint32_t& a,
                                           uint32_t& s) {
    static_assert((SLOTS & (SLOTS - 1)) == 0, "wide probing needs a power-of-two table");
    uint32_t h = ((a - tb) >> 2) + 1u;  // first slot of the window
    // the table always keeps empty slots (<= 4091 entries); the bound only keeps a corrupted
    // table from hanging the warp
    for (int round = 0; round < SLOTS / kWarpSize + 1; round++) {
        const uint32_t sa = tb + 4u * ((h + (uint32_t)lane) & (uint32_t)(SLOTS - 1));
        const uint32_t v = lds_u32(sa);
        const uint32_t bm = __ballot_sync(kFullMask, (v & 0xFFFFFu) == key && v != 0u);
        const uint32_t be = __ballot_sync(kFullMask, v == 0u);
        const uint32_t stop = bm | be;
        if (stop) {
            const int pos = __ffs(stop) - 1;
            a = tb + 4u * ((h + (uint32_t)pos) & (uint32_t)(SLOTS - 1));
            s = __shfl_sync(kFullMask, v, pos);
            return (bm >> pos) & 1u;
        }
        h += kWarpSize;
    }
    s = 0u;
    return false;
}

// The match loop of encoder.rs:313-337 / 639-651 over one staged tile, executed by every lane of
// the warp with identical values (warp-uniform control flow: a diverged warp pays ~20 cycles
// per branch, profiles/r01_encode_ncu.md).
//
// ncu on the first versions showed the loop is bound by instruction issue of a single warp
// (about 6 cycles per issued instruction, profiles/r01_encode_v2_ncu.txt), not by shared-memory
// latency, so the loop is written for the fewest instructions per input byte: one probe per
// byte, the prefix carried pre-shifted (`ph` = code << 8, so key = ph | byte), four bytes per
// trip, everything that is not hit / clean miss out of line.
// GUARD adds the `&mut [u8]`-writer capacity check; it is only instantiated for tiles that
// could overflow the output slot (`room` = bits the writer still accepts).
template <int SLOTS, bool CHECK, bool FIXED, bool GUARD>
__device__ __forceinline__ uint32_t match_tile(uint32_t* __restrict__ table,
                                               uint16_t* __restrict__ codes, const int lane,
                                               const uint8_t* __restrict__ t, uint32_t& i_io,
                                               const uint32_t len, MatchState& m, int32_t room,
                                               const uint32_t max_code, const uint32_t first_code,
                                               const uint32_t clear_code, const uint32_t cs,
                                               const uint32_t inc) {
    if (i_io >= len) return R_TILE_END;
    const uint32_t tb = (uint32_t)__cvta_generic_to_shared(table);
    const uint32_t cbase = (uint32_t)__cvta_generic_to_shared(codes);
    const uint32_t t0 = (uint32_t)__cvta_generic_to_shared(t);
    uint32_t tp = t0 + i_io;         // shared address of the next input byte
    const uint32_t te = t0 + len;    // end of the tile
    uint32_t cp = cbase + 2u * m.ncodes;
    uint32_t ph = (m.pw >> 20) << 8;  // current_prefix, pre-shifted into key position
    uint32_t nh = m.next_code << 20;  // tree.len(), pre-shifted into slot position
    uint32_t ws = FIXED ? 12u : m.write_size;
    uint32_t wtag = ws << 12;
    uint32_t mask = m.mask;
    // inserts left before something happens: width bump / reset (variable), table full (fixed)
    uint32_t until = FIXED ? 4096u - m.next_code : mask - m.next_code + 1u;
    uint32_t reason = R_TILE_END;

    // One byte of encoder.rs:313-337.  `continue`-style flow is done with gotos so that the
    // four unrolled copies share the out-of-line blocks' code shape.
#define SLZW_STEP(OFF, K)                                                                       \
    {                                                                                           \
        const uint32_t k = (K);                                                                 \
        if (CHECK && k > max_code) { /* encoder.rs:315-317 */                                   \
            m.status = SLZW_ERR_UNEXPECTED_CODE;                                                \
            m.detail = k;                                                                       \
            reason = R_STOP;                                                                    \
            tp += OFF;                                                                          \
            goto done;                                                                          \
        }                                                                                       \
        const uint32_t key = ph | k;                                                            \
        uint32_t a = tb + 4u * slot_of_key<SLOTS>(key);                                         \
        uint32_t s = lds_u32(a);                                                                \
        if ((s & 0xFFFFFu) == key && s != 0u) { /* find_word hit, encoder.rs:319-320 */         \
            ph = (s >> 12) & 0xFFF00u;                                                          \
        } else {                                                                                \
            bool hit = false;                                                                   \
            if (s != 0u) hit = probe_wide<SLOTS>(tb, key, lane, a, s);                          \
            if (hit) {                                                                          \
                ph = (s >> 12) & 0xFFF00u;                                                      \
            } else {                                                                            \
                /* miss: encoder.rs:322-324 / 645-649 */                                        \
                sts_u16(cp, (ph >> 8) | wtag);                                                  \
                cp += 2u;                                                                       \
                ph = k << 8;                                                                    \
                if (GUARD) room -= (int32_t)ws;                                                 \
                if (!FIXED || until != 0u) {                                                    \
                    sts_u32(a, nh | key);                                                       \
                    nh += 1u << 20;                                                             \
                    until--;                                                                    \
                    if (!FIXED && until == 0u) { /* new index == mask, encoder.rs:326 */        \
                        tp += OFF + 1;                                                          \
                        goto bump;                                                              \
                    }                                                                           \
                }                                                                               \
                if (GUARD && room < 0) {                                                        \
                    tp += OFF + 1;                                                              \
                    goto full;                                                                  \
                }                                                                               \
            }                                                                                   \
        }                                                                                       \
    }

    for (;;) {
        while (tp + 4u <= te) {
            // the four input bytes are fetched up front so their latency is off the chain
            const uint32_t b0 = lds_u8(tp), b1 = lds_u8(tp + 1), b2 = lds_u8(tp + 2),
                           b3 = lds_u8(tp + 3);
            SLZW_STEP(0, b0)
            SLZW_STEP(1, b1)
            SLZW_STEP(2, b2)
            SLZW_STEP(3, b3)
            tp += 4u;
        }
        while (tp < te) {
            SLZW_STEP(0, lds_u8(tp))
            tp += 1u;
        }
        break;
    bump:
        if (GUARD && room < 0) goto full;
        if (ws < 12u) {  // encoder.rs:327-328
            ws++;
            wtag = ws << 12;
            mask = (1u << ws) - inc;
            until = mask - (nh >> 20) + 1u;
            continue;
        }
        // encoder.rs:329-333: clear code at 12 bits, dictionary restarts
        sts_u16(cp, clear_code | (12u << 12));
        cp += 2u;
        ws = cs + 1;
        mask = (1u << ws) - inc;
        nh = first_code << 20;
        until = mask - first_code + 1u;
        if (GUARD && room - 12 < 0) goto full_noinc;
        reason = R_RESET;
        break;
    full:  // the writer is full (io.rs:244 / 307)
    full_noinc:
        m.status = SLZW_ERR_IO_WRITE_ZERO;
        reason = R_STOP;
        break;
    }
done:
#undef SLZW_STEP
    m.ncodes = (cp - cbase) >> 1;
    // nh wraps at code 4096 (reachable for one step with the default strategy), so the count
    // is derived from `until`, which is exact
    m.next_code = FIXED ? 4096u - until : mask + 1u - until;
    m.write_size = ws;
    m.mask = mask;
    m.pw = (ph >> 8) << 20;
    i_io = tp - t0;
    return reason;
}

template <int SLOTS, int TILE>
__device__ void encode_stream(const DevBatch& a, uint32_t sid, EncWarpSmem<SLOTS, TILE>& S,
                              int lane) {
    using Smem = EncWarpSmem<SLOTS, TILE>;
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
    const bool fixed = a.p.flavour == SLZW_FLAVOUR_FIXED;
    const bool big = a.p.big_endian != 0;
    const uint32_t inc = (!fixed && a.p.tiff_early_change) ? 1u : 0u;
    uint32_t cs = fixed ? 8u : (a.code_size ? a.code_size[sid] : a.p.code_size);

    if (!fixed && (cs < 2 || cs > 8)) {  // encoder.rs:281-283
        if (lane == 0) {
            a.out_len[sid] = 0;
            a.status[sid] = SLZW_ERR_CODE_SIZE;
            a.detail[sid] = cs;
        }
        return;
    }

    // first input tile in flight while the dictionary is cleared
    uint64_t pos = n > 0 ? 1 : 0;  // the first byte is consumed below (encoder.rs:311)
    uint32_t tile_len = (uint32_t)((n - pos) < (uint64_t)TILE ? (n - pos) : TILE);
    uint32_t skew = 0;
    int buf = 0;
    if (tile_len) skew = stage_tile_async(src + pos, tile_len, S.tile[0], lane);

    // cooperative clear of the dictionary and the packed window
    for (int i = lane; i < SLOTS / 4; i += kWarpSize)
        reinterpret_cast<uint4*>(S.table)[i] = make_uint4(0, 0, 0, 0);
    for (int i = lane; i < Smem::kOutWords; i += kWarpSize) S.outw[i] = 0;

    const uint32_t max_code = (1u << cs) - 1;  // encoder.rs:285
    const uint32_t clear_code = 1u << cs;      // encoder.rs:290
    const uint32_t eoi = clear_code + 1;       // encoder.rs:291
    const uint32_t first_code = fixed ? 256u : clear_code + 2;

    MatchState m;
    m.pw = 0;
    m.next_code = first_code;
    m.write_size = fixed ? 12u : cs + 1;
    m.mask = (1u << m.write_size) - inc;
    m.ncodes = 0;
    m.status = SLZW_OK;
    m.detail = 0;

    // ---- packed-window state (warp-uniform) ----
    const uint32_t mis = dst ? (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 3u) : 0u;
    uint32_t qbits = 8 * mis;  // bit cursor inside the window
    uint64_t wbase = 0;        // aligned words already flushed
    uint64_t bits = 0;         // bits handed to the bit writer so far
    // the writer fails when byte index `cap` is written: (bits >> 3) > cap  <=>  bits > limit
    const uint64_t limit = cap > (~0ull - 7) / 8 ? ~0ull : cap * 8 + 7;

    auto push = [&](uint32_t code, uint32_t width) {  // BitWriter::write, io.rs:234-237, 296-300
        S.codes[m.ncodes++] = (uint16_t)((code & ((1u << width) - 1)) | (width << 12));
    };

    // Whole warp: bit-pack the buffered codes, flush complete words.
    auto pack_and_flush = [&](uint32_t count) {
        for (uint32_t base = 0; base < count; base += kWarpSize) {
            const uint32_t idx = base + lane;
            const uint32_t e = idx < count ? S.codes[idx] : 0u;
            const uint32_t w = e >> 12;
            const uint32_t code = e & 0xFFFu;
            uint32_t x = w;
#pragma unroll
            for (int d = 1; d < kWarpSize; d <<= 1) {
                const uint32_t y = __shfl_up_sync(kFullMask, x, d);
                if (lane >= d) x += y;
            }
            const uint32_t off = qbits + x - w;
            const uint32_t total = __shfl_sync(kFullMask, x, kWarpSize - 1);
            if (w) {
                const uint32_t wi = off >> 5, sh = off & 31u;
                if (!big) {
                    const uint64_t v = (uint64_t)code << sh;
                    atomicOr(&S.outw[wi], (uint32_t)v);
                    if (v >> 32) atomicOr(&S.outw[wi + 1], (uint32_t)(v >> 32));
                } else {
                    const uint64_t v = (uint64_t)code << (64 - w - sh);
                    atomicOr(&S.outw[wi], (uint32_t)(v >> 32));
                    if ((uint32_t)v) atomicOr(&S.outw[wi + 1], (uint32_t)v);
                }
            }
            qbits += total;
            bits += total;
        }
        __syncwarp();
        const uint32_t cw = qbits >> 5;
        if (dst)
            for (uint32_t w = lane; w < cw; w += kWarpSize)
                store_word(dst, mis, cap, wbase + w, S.outw[w], big);
        const uint32_t carry = S.outw[cw];
        __syncwarp();
        for (uint32_t w = lane; w <= cw; w += kWarpSize) S.outw[w] = (w == 0) ? carry : 0u;
        wbase += cw;
        qbits &= 31u;
        __syncwarp();
    };

    __syncwarp();

    {   // executed by every lane with identical values (warp-uniform control flow)
        if (!fixed) push(clear_code, m.write_size);  // encoder.rs:297
        if (n > 0) {
            const uint32_t first = __ldg(src);  // encoder.rs:311 / 637: not range-checked
            m.pw = first << 20;
            if (!fixed && n > 1 && first >= first_code) {
                // find_word would index past tree.nodes (encoder.rs:99) unless the second byte
                // is rejected first (encoder.rs:315-317)
                const uint32_t k = __ldg(src + 1);
                if (k > max_code) {
                    m.status = SLZW_ERR_UNEXPECTED_CODE;
                    m.detail = k;
                } else {
                    m.status = SLZW_ERR_REFERENCE_PANIC;
                }
            }
        }
    }
    // the leading clear code alone overflows a tiny slot before the first byte is even read
    if (!fixed && (uint64_t)m.write_size > limit) m.status = SLZW_ERR_IO_WRITE_ZERO;
    uint32_t status = __shfl_sync(kFullMask, m.status, 0);

    while (status == SLZW_OK && pos < n) {
        // prefetch the next tile into the other buffer, then wait for the current one
        const uint64_t npos = pos + tile_len;
        const uint32_t nlen = (uint32_t)((n - npos) < (uint64_t)TILE ? (n - npos) : TILE);
        uint32_t nskew = 0;
        if (nlen) {
            nskew = stage_tile_async(src + npos, nlen, S.tile[buf ^ 1], lane);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
        // The guarded loop is only needed when this tile could overflow the slot (upper bound:
        // 12 bits per byte plus the codes still buffered).
        const uint64_t pending = (uint64_t)__shfl_sync(kFullMask, m.ncodes, 0) * 12u;
        const uint64_t used = bits + pending;
        const bool guard = (limit > used ? limit - used : 0) < 12ull * (TILE + 8);
        uint32_t i = 0;
        for (;;) {
            uint32_t reason = R_TILE_END;
            {   // every lane runs the match loop redundantly: uniform branches, broadcast loads
                const uint8_t* t = S.tile[buf] + skew;
#define SLZW_MATCH(CHK, FIX, GRD, ROOM)                                                    \
    match_tile<SLOTS, CHK, FIX, GRD>(S.table, S.codes, lane, t, i, tile_len, m, ROOM,   \
                                     max_code, first_code, clear_code, cs, inc)
                if (guard) {
                    // exact room: bits already packed plus the widths of the buffered codes
                    uint32_t pend = 0;
                    for (uint32_t c = 0; c < m.ncodes; c++) pend += S.codes[c] >> 12;
                    const int64_t r = (int64_t)(limit > bits ? limit - bits : 0) - (int64_t)pend;
                    const int32_t room = (int32_t)(r > 0x3FFFFFFF ? 0x3FFFFFFF : r);
                    if (fixed)
                        reason = SLZW_MATCH(false, true, true, room);
                    else
                        reason = SLZW_MATCH(true, false, true, room);
                } else if (fixed) {
                    reason = SLZW_MATCH(false, true, false, 0);
                } else if (cs == 8) {
                    reason = SLZW_MATCH(false, false, false, 0);
                } else {
                    reason = SLZW_MATCH(true, false, false, 0);
                }
#undef SLZW_MATCH
            }
            reason = __shfl_sync(kFullMask, reason, 0);
            if (reason == R_RESET) {  // tree.reset(), encoder.rs:332
                for (int j = lane; j < SLOTS / 4; j += kWarpSize)
                    reinterpret_cast<uint4*>(S.table)[j] = make_uint4(0, 0, 0, 0);
                __syncwarp();
                continue;
            }
            break;
        }
        status = __shfl_sync(kFullMask, m.status, 0);
        const uint32_t cnt = __shfl_sync(kFullMask, m.ncodes, 0);
        pack_and_flush(cnt);
        m.ncodes = 0;
        pos = npos;
        tile_len = nlen;
        skew = nskew;
        buf ^= 1;
    }
    cp_async_wait<0>();

    // tail codes (only when the whole input was consumed), then pack whatever is buffered --
    // on an error the codes written before it stay in the output, like the reference's writer
    if (status == SLZW_OK) {
        if (n > 0) push(m.pw >> 20, m.write_size);  // encoder.rs:339 / 653
        if (!fixed) push(eoi, m.write_size);        // encoder.rs:303 / 340
    }
    {
        const uint32_t cnt = __shfl_sync(kFullMask, m.ncodes, 0);
        pack_and_flush(cnt);
        m.ncodes = 0;
    }
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
            uint32_t v = S.outw[0];
            if (big) v = __byte_perm(v, 0, 0x0123);
            for (int j = 0; j < 4; j++) {
                const int64_t b = b0 + j;
                if (b >= 0 && (uint64_t)b < total) dst[b] = (uint8_t)(v >> (8 * j));
            }
        }
        a.out_len[sid] = total;
        a.status[sid] = status;
        a.detail[sid] = m.detail;
    }
    __syncwarp();
}

template <int SLOTS, int TILE, int WARPS>
__global__ void __launch_bounds__(WARPS * kWarpSize, 1) slzw_encode_kernel(const DevBatch a) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    using Smem = EncWarpSmem<SLOTS, TILE>;
    const int warp = threadIdx.x / kWarpSize;
    const int lane = threadIdx.x % kWarpSize;
    Smem& S = reinterpret_cast<Smem*>(smem_raw)[warp];
    for (;;) {
        unsigned long long q = 0;
        if (lane == 0) q = atomicAdd(a.queue, 1ull);
        q = __shfl_sync(kFullMask, q, 0);
        if (q >= a.n) break;
        const uint32_t sid = a.order ? a.order[q] : (uint32_t)q;
        encode_stream<SLOTS, TILE>(a, sid, S, lane);
    }
}

}  // namespace

// ---- launch configuration ---------------------------------------------------------------------
// {dictionary slots, input tile, warps per CTA}: 4096 slots = 16 KB per stream (load <= 0.94,
// wide probing) lets 12-13 streams share one SM's shared memory.
template <int SLOTS, int TILE, int WARPS>
struct EncConfig {
    static size_t smem() { return sizeof(EncWarpSmem<SLOTS, TILE>) * WARPS; }
    static cudaError_t configure() {
        return cudaFuncSetAttribute(slzw_encode_kernel<SLOTS, TILE, WARPS>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem());
    }
    static cudaError_t launch(const DevBatch& a, int grid, cudaStream_t stream) {
        slzw_encode_kernel<SLOTS, TILE, WARPS><<<grid, WARPS * kWarpSize, smem(), stream>>>(a);
        return cudaGetLastError();
    }
};

using Enc0 = EncConfig<4096, 192, 13>;  // default: measured fastest (profiles/r01_encode_configs.md)
using Enc1 = EncConfig<4096, 256, 12>;
using Enc2 = EncConfig<8192, 512, 6>;
using Enc3 = EncConfig<4096, 256, 8>;

static int g_enc_config = 0;

void encode_select_config(int c) { g_enc_config = (c >= 0 && c <= 3) ? c : 0; }
int encode_warps_per_cta() {
    switch (g_enc_config) {
        case 1: return 12;
        case 2: return 6;
        case 3: return 8;
        default: return 13;
    }
}

cudaError_t encode_configure() {
    cudaError_t e;
    if ((e = Enc0::configure()) != cudaSuccess) return e;
    if ((e = Enc1::configure()) != cudaSuccess) return e;
    if ((e = Enc2::configure()) != cudaSuccess) return e;
    return Enc3::configure();
}

cudaError_t encode_launch(const DevBatch& a, int grid, cudaStream_t stream) {
    switch (g_enc_config) {
        case 1: return Enc1::launch(a, grid, stream);
        case 2: return Enc2::launch(a, grid, stream);
        case 3: return Enc3::launch(a, grid, stream);
        default: return Enc0::launch(a, grid, stream);
    }
}

}  // namespace slzw
