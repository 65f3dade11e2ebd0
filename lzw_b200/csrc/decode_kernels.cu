// decode_kernels.cu -- batched LZW decoder for sm_100a.
//
// slzw_decode_exact_kernel: one warp per stream, persistent CTAs, streams handed out through
// the scheduler's work queue.  It reproduces VariableDecoder::inner_decode (decoder.rs:174-290)
// and FixedDecoder::inner_decode (decoder.rs:553-642) state for state -- prefix/suffix/length
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
                    } else if (!fixed) {
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
    for (;;) {
        unsigned long long q = 0;
        if (lane == 0) q = atomicAdd(a.queue, 1ull);
        q = __shfl_sync(kFullMask, q, 0);
        if (q >= a.n) break;
        const uint32_t sid = a.order ? a.order[q] : (uint32_t)q;
        decode_stream_exact<TILE, OUTB>(a, sid, S, lane);
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

}  // namespace slzw
