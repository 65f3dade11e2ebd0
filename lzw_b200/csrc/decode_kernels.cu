// decode_kernels.cu -- batched LZW decoder for sm_100a.
//
// Two kernels, both one warp per stream, persistent CTAs, streams handed out through the
// scheduler's work queue:
//
// slzw_decode_fast_kernel -- the throughput path (32 warps per SM: 12 with their table in shared
// memory, 20 with it in an L2-resident block of global memory).  The reference's prefix-chain walk
// (decoder.rs:251-267) is replaced by an (output offset, length) table: entry n is the word that
// was written for the previous code plus the first byte of the current word, and those bytes are
// contiguous in the stream's own output, so an entry is just (offset of the previous word,
// its length + 1) (equivalent to decoder.rs:272-275).  The warp decodes up to 32 codes per
// step, one per lane: parallel bit extraction (io.rs:43-55 / 113-128; all codes of a step have
// the same width because steps end where the width changes, decoder.rs:277-280), table lookup,
// resolution of codes that name entries created inside the same step (including the KwKwK case,
// decoder.rs:244-250), a prefix sum of the word lengths, then a byte-parallel gather: every
// output byte of the step finds its word through a bit mask of word starts, follows source
// pointers that still point into the step itself, loads its byte from the stream's already
// decoded output and stores it, coalesced.  Results identical to the reference are produced in the
// kernel for success, input that ends before the end-of-information code (Io(UnexpectedEof)), a
// code beyond the table (UnexpectedCode) and a full output slot (Io(WriteZero)).  Everything else
// -- a missing clear code, a first code after a clear that is not a root (decoder.rs:230-236), a
// word longer than the reference's stack, more than 1 MiB of output between two clear codes
// (entries hold 20-bit offsets relative to the last clear), a slot of 1 GiB or more -- makes the
// kernel DEFER the stream: it is appended to a retry list and decoded again, from scratch, by the
// exact kernel.
//
// slzw_decode_exact_kernel -- reproduces VariableDecoder::inner_decode (decoder.rs:174-290) and
// FixedDecoder::inner_decode (decoder.rs:553-642) state for state -- prefix/suffix/length
// tables and the word stack live in shared memory -- so that every observable result matches
// the reference, including the ones that depend on stale table contents (tables are not
// cleared on a clear code, decoder.rs:222-227; the first code after a clear is not
// range-checked, decoder.rs:230-236).  The warp stages compressed tiles into shared memory and
// drains the decoded bytes from a shared-memory window with coalesced stores; lane 0 runs the
// code loop.
#include "slzw_device.cuh"

namespace slzw {

namespace {

constexpr int kMaxTable = 4096;  // decoder.rs:185
constexpr int kMaxStack = 4091;  // decoder.rs:192

enum DecReason : uint32_t { D_DONE = 0, D_REFILL = 1, D_FLUSH = 2 };

template <int TILE, int OUTB>
struct DecWarpSmem {
    uint16_t prefix[kMaxTable];
    uint16_t length[kMaxTable];  // saturating: any value > kMaxStack behaves alike (see below)
    uint8_t suffix[kMaxTable];
    uint8_t stack[kMaxTable];    // decoding_stack (4091 used)
    __align__(16) uint8_t outbuf[OUTB];
    __align__(16) uint8_t tile[TILE + 16];
};

template <int TILE, int OUTB>
__device__ void decode_stream_exact(const DevBatch& a, uint32_t sid, DecWarpSmem<TILE, OUTB>& S,
                                    int lane) {
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
    const bool lenient = a.p.flavour == SLZW_FLAVOUR_VARIABLE_LENIENT;  // full table freezes, no error
    const bool big = a.p.big_endian != 0;
    const uint32_t inc = (!fixed && a.p.tiff_early_change) ? 1u : 0u;
    const uint32_t cs = fixed ? 8u : (a.code_size ? a.code_size[sid] : a.p.code_size);

    if (!fixed && (cs < 2 || cs > 8)) {  // decoder.rs:180-182
        if (lane == 0) {
            a.out_len[sid] = 0;
            a.status[sid] = SLZW_ERR_CODE_SIZE;
            a.detail[sid] = cs;
        }
        return;
    }

    // decoder.rs:197-206 / 569-578: zeroed tables, roots prefilled
    {
        uint4* z0 = reinterpret_cast<uint4*>(S.prefix);
        constexpr int kZero16 = (int)((sizeof(S.prefix) + sizeof(S.length) + sizeof(S.suffix) +
                                       sizeof(S.stack)) / 16);
        for (int i = lane; i < kZero16; i += kWarpSize) z0[i] = make_uint4(0, 0, 0, 0);
        __syncwarp();
        const uint32_t roots = 1u << cs;
        for (uint32_t c = lane; c < roots; c += kWarpSize) {
            S.suffix[c] = (uint8_t)c;
            S.length[c] = 1;
        }
        __syncwarp();
    }

    // ---- lane-0 decoder state (names follow decoder.rs:208-217) ----
    const uint32_t clear_code = fixed ? 256u : (1u << cs);  // fixed: root limit (decoder.rs:614)
    const uint32_t eoi = clear_code + 1;
    uint32_t read_size = fixed ? 12u : cs + 1;
    uint32_t mask = (1u << read_size) - inc;
    uint32_t next_index = fixed ? 256u : clear_code + 2;
    bool have_prev = false;
    uint32_t previous_code = 0, initial_code = 0;
    uint32_t word_length = 0;
    uint32_t byte_buffer = 0, cursor = 0;  // io.rs:16-17 / 86-87
    uint32_t status = SLZW_OK, detail = 0;
    uint64_t produced = 0;  // bytes handed to the writer
    // resumable word copy
    bool copying = false;
    uint32_t j = 0, wl_eff = 0;
    uint32_t ob = 0;  // bytes in outbuf

    uint64_t pos = 0;      // compressed bytes staged so far
    uint64_t flushed = 0;  // decoded bytes already stored to dst
    uint32_t tile_len = 0, tile_i = 0, skew = 0;

    for (;;) {
        uint32_t reason = D_DONE;
        if (lane == 0) {
            const uint8_t* t = S.tile + skew;
            for (;;) {
                if (copying) {
                    if (dst) {
                        while (j < wl_eff) {
                            if (ob == OUTB) break;
                            S.outbuf[ob++] = S.stack[j++];
                        }
                        if (j < wl_eff) {
                            reason = D_FLUSH;
                            break;
                        }
                    }
                    copying = false;
                    if (wl_eff < word_length) {  // write_all on a full slot, decoder.rs:270
                        status = SLZW_ERR_IO_WRITE_ZERO;
                        break;
                    }
                    // decoder.rs:272-284 / 630-636
                    if (next_index < (uint32_t)kMaxTable) {
                        S.prefix[next_index] = (uint16_t)previous_code;
                        S.suffix[next_index] = S.stack[0];
                        const uint32_t l = (uint32_t)S.length[previous_code] + 1u;
                        S.length[next_index] = (uint16_t)(l > 0xFFFFu ? 0xFFFFu : l);
                        next_index++;
                        if (!fixed && next_index == mask && read_size < 12u) {
                            read_size++;
                            mask = (1u << read_size) - inc;
                        }
                    } else if (!fixed && !lenient) {
                        status = SLZW_ERR_MISSING_CLEAR_CODE;
                        break;
                    }
                    previous_code = initial_code;
                }
                // the first-code path below appends one byte without a resumable copy
                if (dst && ob == OUTB) {
                    reason = D_FLUSH;
                    break;
                }
                // BitReader::read_one, io.rs:43-55 / 113-128
                bool eof = false, refill = false;
                while (cursor < read_size) {
                    if (tile_i >= tile_len) {
                        if (pos >= n) eof = true; else refill = true;
                        break;
                    }
                    const uint32_t byte = t[tile_i++];
                    if (!big) byte_buffer |= byte << cursor;
                    else byte_buffer |= byte << (24u - cursor);
                    cursor += 8;
                }
                if (refill) {
                    reason = D_REFILL;
                    break;
                }
                if (eof) {
                    // variable: read_exact fails (decoder.rs:220); fixed: iterator ends
                    // (decoder.rs:585, io.rs:62-64)
                    if (!fixed) status = SLZW_ERR_IO_UNEXPECTED_EOF;
                    break;
                }
                uint32_t code;
                if (!big) {
                    code = byte_buffer & ((1u << read_size) - 1);
                    byte_buffer >>= read_size;
                } else {
                    code = (byte_buffer >> (32u - read_size)) & ((1u << read_size) - 1);
                    byte_buffer <<= read_size;
                }
                cursor -= read_size;

                if (!fixed) {
                    if (code == clear_code) {  // decoder.rs:222-227
                        read_size = cs + 1;
                        mask = (1u << read_size) - inc;
                        next_index = clear_code + 2;
                        have_prev = false;
                        continue;
                    } else if (code == eoi) {  // decoder.rs:228-229
                        break;
                    }
                }
                if (!have_prev) {  // decoder.rs:230-236 / 588-594
                    if (produced < cap) {
                        if (dst) S.outbuf[ob++] = S.suffix[code];
                        produced++;
                    } else {
                        status = SLZW_ERR_IO_WRITE_ZERO;
                        break;
                    }
                    have_prev = true;
                    previous_code = code;
                    S.stack[0] = (uint8_t)code;
                    word_length = 1;
                    continue;
                }
                initial_code = code;
                if (code > next_index) {  // decoder.rs:241-243
                    status = SLZW_ERR_UNEXPECTED_CODE;
                    detail = code;
                    break;
                } else if (code == next_index) {  // decoder.rs:244-250
                    if (word_length >= (uint32_t)kMaxStack) {
                        status = SLZW_ERR_REFERENCE_PANIC;
                        break;
                    }
                    S.stack[word_length] = S.stack[0];
                    word_length++;
                } else {  // decoder.rs:251-267
                    word_length = S.length[code];
                    uint32_t stack_top = word_length;
                    bool bad = false;
                    while (code >= clear_code) {
                        stack_top--;
                        if (stack_top == 0) {  // decoder.rs:258-260
                            status = SLZW_ERR_UNEXPECTED_CODE;
                            detail = code;
                            bad = true;
                            break;
                        }
                        if (stack_top >= (uint32_t)kMaxStack) {  // index panic, decoder.rs:262
                            status = SLZW_ERR_REFERENCE_PANIC;
                            bad = true;
                            break;
                        }
                        S.stack[stack_top] = S.suffix[code];
                        code = S.prefix[code];
                    }
                    if (bad) break;
                    S.stack[0] = (uint8_t)code;  // decoder.rs:266
                }
                if (word_length > (uint32_t)kMaxStack) {  // slice panic, decoder.rs:270
                    status = SLZW_ERR_REFERENCE_PANIC;
                    break;
                }
                {
                    const uint64_t room = cap - produced;
                    wl_eff = (uint64_t)word_length <= room ? word_length : (uint32_t)room;
                    produced += wl_eff;
                    j = 0;
                    copying = true;
                }
            }
        }
        reason = __shfl_sync(kFullMask, reason, 0);
        const uint32_t nb = __shfl_sync(kFullMask, ob, 0);
        if (reason == D_FLUSH || reason == D_DONE) {
            if (dst) {
                for (uint32_t b = lane; b < nb; b += kWarpSize) dst[flushed + b] = S.outbuf[b];
                flushed += nb;
            }
            ob = 0;
            __syncwarp();
            if (reason == D_DONE) break;
        } else {  // D_REFILL
            tile_len = (uint32_t)((n - pos) < (uint64_t)TILE ? (n - pos) : TILE);
            skew = stage_tile(src + pos, tile_len, S.tile, lane);
            pos += tile_len;
            tile_i = 0;
            __syncwarp();
        }
    }

    if (lane == 0) {
        a.out_len[sid] = produced;
        a.status[sid] = status;
        a.detail[sid] = detail;
    }
    __syncwarp();
}

template <int TILE, int OUTB, int WARPS>
__global__ void __launch_bounds__(WARPS * kWarpSize, 1) slzw_decode_exact_kernel(const DevBatch a) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    using Smem = DecWarpSmem<TILE, OUTB>;
    const int warp = threadIdx.x / kWarpSize;
    const int lane = threadIdx.x % kWarpSize;
    Smem& S = reinterpret_cast<Smem*>(smem_raw)[warp];
    // n_dev != nullptr: the streams are the ones the fast kernel deferred (count on the device)
    const uint64_t count = a.n_dev ? (uint64_t)*reinterpret_cast<const volatile uint32_t*>(a.n_dev) : a.n;
    for (;;) {
        unsigned long long q = 0;
        if (lane == 0) q = atomicAdd(a.queue, 1ull);
        q = __shfl_sync(kFullMask, q, 0);
        if (q >= count) break;
        const uint32_t sid = a.order ? a.order[q] : (uint32_t)q;
        decode_stream_exact<TILE, OUTB>(a, sid, S, lane);
    }
}


// ---- fast kernel ---------------------------------------------------------------------------------
// Gather load of one already decoded byte.  A volatile asm statement: the compiler may not sink it
// below the stores that follow (it cannot see that they never alias).
__device__ __forceinline__ uint8_t ld_byte(const uint8_t* p) {
    uint32_t v;
    asm volatile("ld.global.u8 %0, [%1];\n" : "=r"(v) : "l"(p));
    return (uint8_t)v;
}

// Table accesses of the warps whose table lives in global memory: the 47 MB of tables are meant to
// stay in L2 while 4.4 GB of input and output stream through it.  Without a hint the streaming
// lines push table lines out, dirty: the round-1 capture at 65,536 strips shows 4.71 GB of DRAM
// writes for 2.41 GB of output (profiles/r02_decode_65536_ncu_summary.txt).  An evict_last policy on
// every table access keeps them (createpolicy + L2::cache_hint).
__device__ __forceinline__ uint64_t l2_keep_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint32_t table_ld(const uint32_t* p, bool global_table, uint64_t pol) {
    if (!global_table) return *p;
    uint32_t v;
    asm volatile("ld.global.L2::cache_hint.u32 %0, [%1], %2;\n" : "=r"(v) : "l"(p), "l"(pol) : "memory");
    return v;
}
__device__ __forceinline__ void table_st(uint32_t* p, uint32_t v, bool global_table, uint64_t pol) {
    if (!global_table) {
        *p = v;
        return;
    }
    asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;\n" ::"l"(p), "r"(v), "l"(pol) : "memory");
}

constexpr int kFastWin = 1024;   // output bytes one step may produce (32 mask words)
constexpr int kFastTile = 512;   // compressed bytes staged per tile
constexpr uint32_t kLit = 0x80000000u;
// Table entries hold 20-bit output offsets relative to the output position of the last clear
// code (entries never outlive a clear): a dictionary generation may span 1 MiB of output.
constexpr uint32_t kFastMaxSpan = 1u << 20;

// Per-warp working set besides the table.  The table itself (4096 entries = offset << 12 | length)
// lives in shared memory for the first SW warps of a CTA and in global memory (a per-warp 16 KB
// block of the context's scratch, L2-resident: 148 SMs x GW x 16 KB) for the other GW warps: the
// kernel is latency-bound with the 13 warps per SM that shared memory has room for, and a table
// lookup happens once per 32 codes, so paying L2 latency for it is cheap next to doubling the
// number of streams in flight.
struct FastWarpSmem {
    uint32_t bits[kFastWin / 32];        // word starts of the current step
    __align__(16) uint8_t tile[kFastTile + 32];
};

// Returns false when the stream has to be decoded by the exact kernel.
__device__ bool decode_stream_fast(const DevBatch& a, uint32_t sid, uint32_t* __restrict__ table,
                                   const bool global_table, const uint64_t pol, FastWarpSmem& S, int lane) {
    const uint64_t in_begin = a.in_off[sid];
    const uint64_t n = a.in_off[sid + 1] - in_begin;
    const uint8_t* src = a.in + in_begin;
    uint8_t* dst = nullptr;
    uint64_t cap = ~0ull;
    if (a.out != nullptr) {
        const uint64_t ob = a.out_off[sid];
        dst = a.out + ob;
        cap = a.out_off[sid + 1] - ob;
        if (cap >= 0x80000000ull) return false;  // offsets are 31 bits + literal flag
    }
    const bool fixed = a.p.flavour == SLZW_FLAVOUR_FIXED;
    const bool lenient = a.p.flavour == SLZW_FLAVOUR_VARIABLE_LENIENT;  // full table freezes, no error
    const bool big = a.p.big_endian != 0;
    const uint32_t inc = (!fixed && a.p.tiff_early_change) ? 1u : 0u;
    const uint32_t cs = fixed ? 8u : (a.code_size ? a.code_size[sid] : a.p.code_size);
    if (!fixed && (cs < 2 || cs > 8)) {  // decoder.rs:180-182
        if (lane == 0) {
            a.out_len[sid] = 0;
            a.status[sid] = SLZW_ERR_CODE_SIZE;
            a.detail[sid] = cs;
        }
        return true;
    }

    const uint32_t roots = fixed ? 256u : (1u << cs);  // codes below are single bytes
    const uint32_t clear_code = roots, eoi = roots + 1;  // variable flavour only
    const uint32_t first_index = fixed ? 256u : roots + 2;
    uint32_t w = fixed ? 12u : cs + 1;       // read_size, decoder.rs:208
    uint32_t mask = (1u << w) - inc;         // decoder.rs:211
    uint32_t nidx = first_index;             // next_index
    bool hp = false;                         // previous_code.is_some()
    uint32_t seg_base = 0;                   // output position of the last clear code
    uint32_t prev_off = 0, prev_len = 0;     // the previous word, in the output
    uint64_t bitpos = 0;
    const uint64_t total_bits = n * 8;
    uint32_t produced = 0;                   // bytes written (<= 2^20 when dst != nullptr)
    uint64_t produced64 = 0;                 // size-only passes are not bounded
    uint32_t status = SLZW_OK, detail = 0;
    uint64_t tile_pos = 0;
    uint32_t tile_len = 0, skew = 0;         // S.tile[skew + j] = src[tile_pos + j], j < tile_len
    const uint32_t lanemask_le = 0xFFFFFFFFu >> (31 - lane);

    for (;;) {
        // ---- how many codes this step may read at the current width ----
        // whole codes left in the input, capped at 32 (no 64-bit division in the common case)
        const uint64_t rem_bits = total_bits - bitpos;
        const uint32_t avail = rem_bits >= 32u * 12u ? 32u : (uint32_t)rem_bits / w;
        if (avail == 0) {
            // fixed: the iterator ends, decoder.rs:585 (leftover bits ignored, io.rs:62-64);
            // variable: read_exact fails, Io(UnexpectedEof), decoder.rs:220 -- the bytes written
            // so far stay (this is how SURVEY.md F1 streams end)
            if (!fixed) status = SLZW_ERR_IO_UNEXPECTED_EOF;
            break;
        }
        uint32_t bmax = avail < 32u ? avail : 32u;
        const uint32_t adj = hp ? 0u : 1u;   // the first code after a clear creates no entry
        if (!((fixed || lenient) && nidx >= (uint32_t)kMaxTable)) {
            const uint32_t room = (w < 12u ? mask : (uint32_t)kMaxTable) - nidx + adj;
            if (room < bmax) bmax = room;
        }
        const uint32_t bn = bmax ? bmax : 1u;  // bmax == 0: table full, a control code must follow

        // ---- stage the compressed bytes of the step ----
        {
            const uint64_t first = bitpos >> 3;
            const uint64_t last = (bitpos + 32u * 12u + 7u) >> 3;  // exclusive upper bound
            if (first < tile_pos || (last > tile_pos + tile_len && tile_pos + tile_len < n)) {
                tile_pos = first;
                tile_len = (uint32_t)((n - tile_pos) < (uint64_t)kFastTile ? (n - tile_pos) : kFastTile);
                __syncwarp();
                skew = stage_tile(src + tile_pos, tile_len, S.tile, lane);
                __syncwarp();
            }
        }

        // ---- one code per lane (io.rs:43-55 / 113-128) ----
        uint32_t c;
        {
            const uint64_t bit = bitpos + (uint64_t)lane * w;
            const uint32_t ba = skew + (uint32_t)((bit >> 3) - tile_pos);
            const uint32_t* tw = reinterpret_cast<const uint32_t*>(S.tile);
            uint32_t v = 0;
            if ((uint32_t)lane < bn) v = __funnelshift_r(tw[ba >> 2], tw[(ba >> 2) + 1], (ba & 3u) * 8u);
            const uint32_t sh = (uint32_t)bit & 7u;
            if (!big) c = (v >> sh) & ((1u << w) - 1u);
            else c = (__byte_perm(v, 0, 0x0123) >> (32u - w - sh)) & ((1u << w) - 1u);
        }
        // control codes end the step (decoder.rs:222-229)
        uint32_t b = bn;
        uint32_t ctrl = 0xFFFFFFFFu;  // the control code that ends this step, if any
        if (!fixed) {
            const uint32_t cm = __ballot_sync(kFullMask, (uint32_t)lane < bn && (c == clear_code || c == eoi));
            if (cm) {
                b = (uint32_t)__ffs(cm) - 1u;
                ctrl = __shfl_sync(kFullMask, c, b);
            }
        }
        if (bmax == 0 && ctrl == 0xFFFFFFFFu) return false;  // MissingClearCode, decoder.rs:281-283

        if (b > 0) {
            // ---- classify (decoder.rs:230-268) ----
            // the entry index this lane's code is compared with (next_index at that moment)
            const uint32_t ni = nidx + (uint32_t)lane - adj;
            // UnexpectedCode (decoder.rs:241-243): the words before it are still written
            {
                const uint32_t um = __ballot_sync(kFullMask, (uint32_t)lane < b && (uint32_t)lane >= adj &&
                                                                 c >= roots && c > ni);
                if (um) {
                    const uint32_t f = (uint32_t)__ffs(um) - 1u;
                    status = SLZW_ERR_UNEXPECTED_CODE;
                    detail = __shfl_sync(kFullMask, c, (int)f);
                    b = f;
                    ctrl = 0xFFFFFFFFu;
                }
            }
        }
        if (b > 0) {
            const bool act = (uint32_t)lane < b;
            bool bad = false;
            uint32_t len = 0, srci = 0;
            int qd = -2;  // >= -1: the word extends the word of step code qd (-1 = previous step)
            if (act) {
                if (c < roots) {
                    len = 1;
                    srci = kLit | c;
                } else if (!hp && lane == 0) {
                    bad = true;  // first code is not a root: stale-table semantics, decoder.rs:230-236
                } else if (c < nidx) {
                    const uint32_t e = table_ld(table + c, global_table, pol);
                    len = e & 0xFFFu;
                    srci = seg_base + (e >> 12);
                } else {  // c <= ni: an entry created inside this step, or the one being created
                    qd = (int)(c - nidx + adj) - 1;
                }
            }
            if (__any_sync(kFullMask, bad)) return false;
            // ---- lengths of words that extend a word of this step ----
            bool known = qd < 0;
            if (qd == -1) len = prev_len + 1u;
            for (;;) {
                if (!__any_sync(kFullMask, !known)) break;
                const int from = qd > 0 ? qd : 0;
                const uint32_t lk = __shfl_sync(kFullMask, len, from);
                const bool kk = __shfl_sync(kFullMask, (int)known, from) != 0;
                if (!known && kk) {
                    len = lk + 1u;
                    known = true;
                }
            }
            // a word longer than the reference's stack panics there (decoder.rs:247, 262, 270)
            if (__any_sync(kFullMask, act && len > (uint32_t)kMaxStack)) return false;
            // ---- output positions ----
            uint32_t incl = act ? len : 0u;
#pragma unroll
            for (int d = 1; d < kWarpSize; d <<= 1) {
                const uint32_t y = __shfl_up_sync(kFullMask, incl, d);
                if (lane >= d) incl += y;
            }
            uint32_t total = __shfl_sync(kFullMask, incl, (int)b - 1);
            if (total > (uint32_t)kFastWin && b > 1) {
                // keep the codes whose words fit the window (at least one)
                const uint32_t fit = __ballot_sync(kFullMask, act && incl <= (uint32_t)kFastWin);
                uint32_t nb = (uint32_t)__popc(fit);
                if (nb == 0) nb = 1;
                b = nb;
                ctrl = 0xFFFFFFFFu;
                status = SLZW_OK;  // a code beyond the table, if any, is met again in a later step
                detail = 0;
                total = __shfl_sync(kFullMask, incl, (int)b - 1);
            }
            const bool act2 = (uint32_t)lane < b;
            const uint32_t pos = incl - len;  // relative to `produced`
            // Io(WriteZero): the `&mut [u8]` writer takes the bytes that still fit, then fails
            // (decoder.rs:231, 270); nothing after that is observable
            const uint32_t nr_set = (total + 31u) >> 5;  // mask words that get word-start bits
            if (dst && (uint64_t)produced + total > cap) {
                status = SLZW_ERR_IO_WRITE_ZERO;
                detail = 0;
                total = (uint32_t)(cap - produced);
            }
            {
                const uint32_t pq = __shfl_sync(kFullMask, pos, qd > 0 ? qd : 0);
                if (qd >= 0) srci = produced + pq;
                else if (qd == -1) srci = prev_off;
            }
            // offsets of this step's entries must fit the entry format
            if (dst && !((fixed || lenient) && nidx >= (uint32_t)kMaxTable) && produced + total - seg_base >= kFastMaxSpan)
                return false;
            // ---- new entries (decoder.rs:272-276 / 630-634) ----
            {
                uint32_t poff = produced + __shfl_up_sync(kFullMask, pos, 1);
                uint32_t plen = __shfl_up_sync(kFullMask, len, 1);
                if (lane == 0) {
                    poff = prev_off;
                    plen = prev_len;
                }
                const uint32_t idx = nidx + (uint32_t)lane - adj;
                if (act2 && (uint32_t)lane >= adj && idx < (uint32_t)kMaxTable)
                    table_st(table + idx, ((poff - seg_base) << 12) | ((plen + 1u) & 0xFFFu), global_table, pol);
            }
            // ---- copy ----
            if (dst) {
                if (nr_set > (uint32_t)(kFastWin / 32)) {
                    // a single long word (b == 1): periodic copy, the source may run into the word
                    const uint32_t so = __shfl_sync(kFullMask, srci, 0);
                    const uint32_t wl = total;
                    if (so & kLit) {
                        if (lane == 0) dst[produced] = (uint8_t)so;
                    } else {
                        const uint32_t dist = produced - so;  // > 0; the word repeats with this period
                        for (uint32_t i = lane; i < wl; i += kWarpSize)
                            dst[produced + i] = dst[so + i % dist];
                    }
                } else {
                    if (act2) atomicOr(&S.bits[pos >> 5], 1u << (pos & 31u));
                    __syncwarp();
                    const uint32_t nr = (total + 31u) >> 5;  // rounds that write bytes
                    const uint32_t wbits = (uint32_t)lane < nr_set ? S.bits[lane] : 0u;
                    uint32_t cnt = (uint32_t)__popc(wbits);  // -> exclusive prefix count of word starts
                    {
                        uint32_t x = cnt;
#pragma unroll
                        for (int d = 1; d < kWarpSize; d <<= 1) {
                            const uint32_t y = __shfl_up_sync(kFullMask, x, d);
                            if (lane >= d) x += y;
                        }
                        cnt = x - cnt;
                    }
                    // Output byte 32 * r + lane of the step: which word it belongs to, where that
                    // word's bytes come from.  Sources that lie inside this step are followed (they
                    // strictly decrease) until they leave it.  Returns the byte, or the output
                    // offset to load it from (kLit clear).
                    auto resolve = [&](uint32_t r, bool& on) -> uint32_t {
                        const uint32_t m = __shfl_sync(kFullMask, wbits, (int)r);
                        const uint32_t cb = __shfl_sync(kFullMask, cnt, (int)r);
                        const uint32_t ob = 32u * r + (uint32_t)lane;
                        on = ob < total;
                        int own = (int)(cb + (uint32_t)__popc(m & lanemask_le)) - 1;
                        if (!on) own = 0;
                        uint32_t sp = __shfl_sync(kFullMask, srci, own);
                        uint32_t i = ob - __shfl_sync(kFullMask, pos, own);
                        for (;;) {
                            const bool inb = on && !(sp & kLit) && sp + i >= produced;
                            if (!__any_sync(kFullMask, inb)) break;
                            const uint32_t rel = inb ? sp + i - produced : 0u;
                            const uint32_t m2 = __shfl_sync(kFullMask, wbits, (int)(rel >> 5));
                            const uint32_t c2 = __shfl_sync(kFullMask, cnt, (int)(rel >> 5));
                            const int own2 = (int)(c2 + (uint32_t)__popc(m2 & (0xFFFFFFFFu >> (31u - (rel & 31u))))) - 1;
                            const uint32_t s2 = __shfl_sync(kFullMask, srci, own2);
                            const uint32_t p2 = __shfl_sync(kFullMask, pos, own2);
                            if (inb) {
                                sp = s2;
                                i = rel - p2;
                            }
                        }
                        return (sp & kLit) ? sp : sp + i;
                    };
                    // two rounds at a time: both gathers are in flight before the first store
                    // (all sources lie below `produced`, the stores at or above it)
                    uint32_t r = 0;
                    for (; r + 2u <= nr; r += 2u) {
                        bool on0, on1;
                        const uint32_t s0 = resolve(r, on0);
                        const uint32_t s1 = resolve(r + 1u, on1);
                        uint8_t v0 = (uint8_t)s0, v1 = (uint8_t)s1;
                        if (on0 && !(s0 & kLit)) v0 = ld_byte(dst + s0);
                        if (on1 && !(s1 & kLit)) v1 = ld_byte(dst + s1);
                        if (on0) dst[produced + 32u * r + (uint32_t)lane] = v0;
                        if (on1) dst[produced + 32u * (r + 1u) + (uint32_t)lane] = v1;
                    }
                    if (r < nr) {
                        bool on0;
                        const uint32_t s0 = resolve(r, on0);
                        if (on0) dst[produced + 32u * r + (uint32_t)lane] = (s0 & kLit) ? (uint8_t)s0 : ld_byte(dst + s0);
                    }
                    if ((uint32_t)lane < nr_set) S.bits[lane] = 0u;
                }
            }
            // ---- state (decoder.rs:272-284) ----
            prev_off = produced + __shfl_sync(kFullMask, pos, (int)b - 1);
            prev_len = __shfl_sync(kFullMask, len, (int)b - 1);
            produced += total;
            produced64 += total;
            bitpos += (uint64_t)b * w;  // all codes of the step were read at the width before a bump
            if (!((fixed || lenient) && nidx >= (uint32_t)kMaxTable)) {
                nidx += b - adj;
                if (!fixed && nidx == mask && w < 12u) {  // decoder.rs:277-280
                    w++;
                    mask = (1u << w) - inc;
                }
            }
            hp = true;
            __syncwarp();  // the step's stores are visible to the loads of later steps
        }
        if (status != SLZW_OK) break;
        if (ctrl != 0xFFFFFFFFu) {
            bitpos += w;
            if (ctrl == eoi) break;  // decoder.rs:228-229, trailing input is ignored
            // decoder.rs:222-227: the tables themselves are left as they are
            w = cs + 1;
            mask = (1u << w) - inc;
            nidx = first_index;
            hp = false;
            seg_base = produced;
        }
    }

    if (lane == 0) {
        a.out_len[sid] = dst ? (uint64_t)produced : produced64;
        a.status[sid] = status;
        a.detail[sid] = detail;
    }
    __syncwarp();
    return true;
}

template <int SW, int GW, int CTAS>
__global__ void __launch_bounds__((SW + GW) * kWarpSize, CTAS) slzw_decode_fast_kernel(const DevBatch a) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int warp = threadIdx.x / kWarpSize;
    const int lane = threadIdx.x % kWarpSize;
    // [SW tables][SW + GW working sets]
    FastWarpSmem& S = reinterpret_cast<FastWarpSmem*>(smem_raw + (size_t)SW * kMaxTable * 4)[warp];
    uint32_t* table = warp < SW ? reinterpret_cast<uint32_t*>(smem_raw) + (size_t)warp * kMaxTable
                                : a.dec_tables + ((size_t)blockIdx.x * GW + (size_t)(warp - SW)) * kMaxTable;
    for (int i = lane; i < kFastWin / 32; i += kWarpSize) S.bits[i] = 0u;
    __syncwarp();
    const bool global_table = warp >= SW;
    const uint64_t pol = l2_keep_policy();
    // a batch smaller than the grid's warps spreads over the SMs: only ceil(n / CTAs) warps per CTA
    // take streams (the first ones, whose tables are in shared memory)
    const uint32_t takers = (uint32_t)((a.n + gridDim.x - 1) / gridDim.x);
    for (; (uint32_t)warp < takers;) {
        unsigned long long q = 0;
        if (lane == 0) q = atomicAdd(a.queue, 1ull);
        q = __shfl_sync(kFullMask, q, 0);
        if (q >= a.n) break;
        const uint32_t sid = a.order ? a.order[q] : (uint32_t)q;
        const bool done = decode_stream_fast(a, sid, table, global_table, pol, S, lane);
        __syncwarp();
        if (!done) {
            // a deferred stream may have left word-start bits behind
            for (int i = lane; i < kFastWin / 32; i += kWarpSize) S.bits[i] = 0u;
            if (lane == 0) a.retry_ids[atomicAdd(a.retry, 1u)] = sid;
            __syncwarp();
        }
    }
}

}  // namespace

constexpr int kDecTile = 512;
constexpr int kDecOutB = 1024;
constexpr int kDecWarps = 8;

size_t decode_exact_smem_bytes() { return sizeof(DecWarpSmem<kDecTile, kDecOutB>) * kDecWarps; }
int decode_exact_warps_per_cta() { return kDecWarps; }

cudaError_t decode_exact_configure() {
    return cudaFuncSetAttribute(slzw_decode_exact_kernel<kDecTile, kDecOutB, kDecWarps>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)decode_exact_smem_bytes());
}

cudaError_t decode_exact_launch(const DevBatch& a, int grid, cudaStream_t stream) {
    slzw_decode_exact_kernel<kDecTile, kDecOutB, kDecWarps>
        <<<grid, kDecWarps * kWarpSize, decode_exact_smem_bytes(), stream>>>(a);
    return cudaGetLastError();
}


// {warps with the table in shared memory, warps with the table in global memory}
template <int SW, int GW, int CTAS = 1>
struct FastConfig {
    static size_t smem() { return (size_t)SW * kMaxTable * 4 + sizeof(FastWarpSmem) * (SW + GW); }
    static cudaError_t configure() {
        return cudaFuncSetAttribute(slzw_decode_fast_kernel<SW, GW, CTAS>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem());
    }
    static cudaError_t launch(const DevBatch& a, int num_sms, cudaStream_t stream) {
        // one CTA per SM as soon as there is a stream for each (the kernel spreads a small batch)
        const uint64_t cap = (uint64_t)num_sms * CTAS;
        const int grid = (int)(a.n < cap ? a.n : cap);
        slzw_decode_fast_kernel<SW, GW, CTAS><<<grid, (SW + GW) * kWarpSize, smem(), stream>>>(a);
        return cudaGetLastError();
    }
};

// 32 warps per SM saturate the kernel (26: 24.4 ms, 32: 22.2 ms, 40: 23.0 ms, 48: 24.0 ms for
// config 3 at 65,536 strips; 13 with shared-memory tables only: 38.9 ms)
using Fast0 = FastConfig<12, 20>;  // default
using Fast1 = FastConfig<13, 0>;   // shared-memory tables only
using Fast2 = FastConfig<0, 32>;   // global-memory tables only

static int g_fast_config = 0;

void decode_select_config(int c) { g_fast_config = (c >= 0 && c <= 2) ? c : 0; }

// global-memory tables one decode launch needs (bytes)
size_t decode_fast_table_bytes(int num_sms) { return (size_t)num_sms * 32 * kMaxTable * 4; }

cudaError_t decode_fast_configure() {
    cudaError_t e = Fast0::configure();
    if (e != cudaSuccess) return e;
    if ((e = Fast1::configure()) != cudaSuccess) return e;
    return Fast2::configure();
}

cudaError_t decode_fast_launch(const DevBatch& a, int num_sms, cudaStream_t stream) {
    switch (g_fast_config) {
        case 1: return Fast1::launch(a, num_sms, stream);
        case 2: return Fast2::launch(a, num_sms, stream);
        default: return Fast0::launch(a, num_sms, stream);
    }
}

}  // namespace slzw
