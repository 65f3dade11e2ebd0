// Experiment, not part of libslzw.so (compiled only with -DSLZW_EXP_LANES, tools/build_variants.sh):
// one LANE per stream instead of one warp per stream.  Result: profiles/r02_encode_notes.md.
// Included into the anonymous namespace of encode_kernels.cu.
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
__device__ __forceinline__ uint32_t scr(uint32_t code) { return (code * kScr) & 0xFFFu; }
__device__ __forceinline__ uint32_t unscr(uint32_t code) { return (code * kScrInv) & 0xFFFu; }
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

