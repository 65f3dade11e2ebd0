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
//     encoder.rs:313-337) over tiles the whole warp stages into shared memory;
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

template <int SLOTS>
__device__ __forceinline__ uint32_t slot_of(uint32_t key) {
    const uint32_t h = key * 0x9E3779B1u;
    if constexpr ((SLOTS & (SLOTS - 1)) == 0) {
        return h >> (32 - __builtin_ctz(SLOTS));
    } else {
        return __umulhi(h, (uint32_t)SLOTS);
    }
}

enum Reason : uint32_t { R_TILE_END = 0, R_RESET = 1, R_STOP = 2 };

template <int SLOTS, int TILE>
struct EncWarpSmem {
    static constexpr int kCodeBuf = TILE + 16;                  // codes one tile can emit
    static constexpr int kOutWords = (kCodeBuf * 12) / 32 + 4;  // packed window
    uint32_t table[SLOTS];
    uint32_t outw[kOutWords];
    uint16_t codes[kCodeBuf];
    __align__(16) uint8_t tile[TILE + 16];
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

    // cooperative clear of the dictionary and the packed window
    for (int i = lane; i < SLOTS / 4; i += kWarpSize)
        reinterpret_cast<uint4*>(S.table)[i] = make_uint4(0, 0, 0, 0);
    for (int i = lane; i < Smem::kOutWords; i += kWarpSize) S.outw[i] = 0;

    const uint32_t max_code = (1u << cs) - 1;  // encoder.rs:285
    const uint32_t clear_code = 1u << cs;      // encoder.rs:290
    const uint32_t eoi = clear_code + 1;       // encoder.rs:291
    const uint32_t first_code = fixed ? 256u : clear_code + 2;

    // ---- lane-0 match state ----
    uint32_t prefix = 0;
    uint32_t next_code = first_code;                 // tree.len()
    uint32_t write_size = fixed ? 12u : cs + 1;      // encoder.rs:289
    uint32_t mask = (1u << write_size) - inc;        // encoder.rs:292
    uint64_t bits = 0;                               // bits handed to the bit writer so far
    uint32_t status = SLZW_OK, detail = 0;
    uint32_t ncodes = 0;

    // ---- packed-window state (warp-uniform) ----
    const uint32_t mis = dst ? (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 3u) : 0u;
    uint32_t qbits = 8 * mis;  // bit cursor inside the window
    uint64_t wbase = 0;        // aligned words already flushed

    auto push = [&](uint32_t code, uint32_t width) {  // BitWriter::write, io.rs:234-237, 296-300
        S.codes[ncodes++] = (uint16_t)((code & ((1u << width) - 1)) | (width << 12));
        bits += width;
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
                const uint32_t wi = off >> 5, s = off & 31u;
                if (!big) {
                    const uint64_t v = (uint64_t)code << s;
                    atomicOr(&S.outw[wi], (uint32_t)v);
                    if (v >> 32) atomicOr(&S.outw[wi + 1], (uint32_t)(v >> 32));
                } else {
                    const uint64_t v = (uint64_t)code << (64 - w - s);
                    atomicOr(&S.outw[wi], (uint32_t)(v >> 32));
                    if ((uint32_t)v) atomicOr(&S.outw[wi + 1], (uint32_t)v);
                }
            }
            qbits += total;
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

    bool finished = false;  // stream ended normally: tail codes + fill() pending
    if (lane == 0 && !fixed) {
        push(clear_code, write_size);  // encoder.rs:297
        if ((bits >> 3) > cap) status = SLZW_ERR_IO_WRITE_ZERO;
    }

    uint64_t pos = 0;  // input bytes consumed
    if (n > 0 && lane == 0 && status == SLZW_OK) {
        prefix = __ldg(src);  // encoder.rs:311 / 637: first byte is not range-checked
        if (!fixed && n > 1 && prefix >= first_code) {
            // find_word would index past tree.nodes (encoder.rs:99) unless the second byte is
            // rejected first (encoder.rs:315-317)
            const uint32_t k = __ldg(src + 1);
            if (k > max_code) {
                status = SLZW_ERR_UNEXPECTED_CODE;
                detail = k;
            } else {
                status = SLZW_ERR_REFERENCE_PANIC;
            }
        }
    }
    status = __shfl_sync(kFullMask, status, 0);
    if (n > 0) pos = 1;

    while (status == SLZW_OK && pos < n) {
        const uint32_t tile_len = (uint32_t)((n - pos) < (uint64_t)TILE ? (n - pos) : TILE);
        const uint32_t skew = stage_tile(src + pos, tile_len, S.tile, lane);
        __syncwarp();
        uint32_t i = 0;
        for (;;) {
            uint32_t reason = R_TILE_END;
            if (lane == 0) {
                const uint8_t* t = S.tile + skew;
                while (i < tile_len) {
                    const uint32_t k = t[i++];
                    if (k > max_code) {  // encoder.rs:315-317
                        status = SLZW_ERR_UNEXPECTED_CODE;
                        detail = k;
                        reason = R_STOP;
                        break;
                    }
                    const uint32_t key = (prefix << 8) | k;
                    uint32_t h = slot_of<SLOTS>(key);
                    uint32_t s;
                    while ((s = S.table[h]) != 0u && (s & 0xFFFFFu) != key)
                        h = (h + 1 == SLOTS) ? 0 : h + 1;
                    if (s != 0u) {  // find_word hit, encoder.rs:319-320
                        prefix = s >> 20;
                        continue;
                    }
                    if (fixed) {  // encoder.rs:645-649
                        if (next_code < 4096u) {
                            S.table[h] = (next_code << 20) | key;
                            next_code++;
                        }
                        push(prefix, 12);
                        prefix = k;
                        if ((bits >> 3) > cap) {
                            status = SLZW_ERR_IO_WRITE_ZERO;
                            reason = R_STOP;
                            break;
                        }
                    } else {  // encoder.rs:322-335
                        const uint32_t idx = next_code++;
                        S.table[h] = (idx << 20) | key;
                        push(prefix, write_size);
                        prefix = k;
                        if ((bits >> 3) > cap) {
                            status = SLZW_ERR_IO_WRITE_ZERO;
                            reason = R_STOP;
                            break;
                        }
                        if (idx == mask) {
                            if (write_size < 12u) {
                                write_size++;
                                mask = (1u << write_size) - inc;
                            } else {
                                push(clear_code, 12);
                                write_size = cs + 1;
                                mask = (1u << write_size) - inc;
                                next_code = first_code;
                                if ((bits >> 3) > cap) {
                                    status = SLZW_ERR_IO_WRITE_ZERO;
                                    reason = R_STOP;
                                } else {
                                    reason = R_RESET;
                                }
                                break;
                            }
                        }
                    }
                }
            }
            reason = __shfl_sync(kFullMask, reason, 0);
            if (reason == R_RESET) {  // tree.reset(), encoder.rs:332
                for (int j = lane; j < SLOTS / 4; j += kWarpSize)
                    reinterpret_cast<uint4*>(S.table)[j] = make_uint4(0, 0, 0, 0);
                __syncwarp();
                i = __shfl_sync(kFullMask, i, 0);
                continue;
            }
            break;
        }
        status = __shfl_sync(kFullMask, status, 0);
        const uint32_t cnt = __shfl_sync(kFullMask, ncodes, 0);
        pack_and_flush(cnt);
        ncodes = 0;
        pos += tile_len;
    }

    // tail codes (only when the whole input was consumed), then pack whatever is buffered --
    // on an error the codes written before it stay in the output, like the reference's writer
    if (status == SLZW_OK && lane == 0) {
        if (n > 0) push(prefix, write_size);  // encoder.rs:339 / 653
        if (!fixed) push(eoi, write_size);    // encoder.rs:303 / 340
        if ((bits >> 3) > cap) status = SLZW_ERR_IO_WRITE_ZERO;
    }
    status = __shfl_sync(kFullMask, status, 0);
    {
        const uint32_t cnt = __shfl_sync(kFullMask, ncodes, 0);
        pack_and_flush(cnt);
        ncodes = 0;
    }
    finished = (status == SLZW_OK);

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
        a.detail[sid] = detail;
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
constexpr int kEncSlots = 6144;  // 24 KB dictionary, load factor <= 0.63
constexpr int kEncTile = 512;
constexpr int kEncWarps = 8;

size_t encode_smem_bytes() { return sizeof(EncWarpSmem<kEncSlots, kEncTile>) * kEncWarps; }
int encode_warps_per_cta() { return kEncWarps; }

cudaError_t encode_configure() {
    return cudaFuncSetAttribute(slzw_encode_kernel<kEncSlots, kEncTile, kEncWarps>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)encode_smem_bytes());
}

cudaError_t encode_launch(const DevBatch& a, int grid, cudaStream_t stream) {
    slzw_encode_kernel<kEncSlots, kEncTile, kEncWarps>
        <<<grid, kEncWarps * kWarpSize, encode_smem_bytes(), stream>>>(a);
    return cudaGetLastError();
}

}  // namespace slzw
