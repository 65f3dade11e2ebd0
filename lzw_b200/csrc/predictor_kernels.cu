// predictor_kernels.cu -- TIFF Predictor = 2 (horizontal differencing, TIFF 6.0 section 14) for
// 8-bit samples, applied in place to every strip of a batch.
//
// Not part of the reference: salzweg stops at the raw code stream (lzw/examples/
// compress_image_data.rs:22-24 sinks it).  SURVEY.md 8f.1 lists the predictor as the container
// step either side of the codec: TIFF writers difference each row before TiffStyleEncoder sees it,
// readers accumulate each row after TiffStyleDecoder.  Both directions are pure HBM traffic (one
// read and one write of every byte), so they work on aligned 16-byte words; the ragged first and
// last word of a strip (difference) or row (accumulate) go byte by byte and never touch a
// neighbouring strip's bytes.
#include "slzw_device.cuh"

namespace slzw {
namespace {

constexpr int kDiffThreads = 256;
constexpr int kAccWarps = 4;
constexpr int kAccLaneWords = 2;  // 16-byte words per lane and pass of the accumulate kernel

struct Word16 {
    uint32_t w[4];
};

// Bytes [jlo, jhi) of the aligned word at p are inside the strip / row; the others read as zero.
__device__ __forceinline__ Word16 load_word(const uint8_t* p, int jlo, int jhi) {
    Word16 v;
    if (jlo == 0 && jhi == 16) {
        const uint4 t = *reinterpret_cast<const uint4*>(p);
        v.w[0] = t.x, v.w[1] = t.y, v.w[2] = t.z, v.w[3] = t.w;
    } else {
        v.w[0] = v.w[1] = v.w[2] = v.w[3] = 0;
#pragma unroll
        for (int j = 0; j < 16; j++)
            if (j >= jlo && j < jhi) v.w[j >> 2] |= (uint32_t)p[j] << (8 * (j & 3));
    }
    return v;
}

__device__ __forceinline__ void store_word(uint8_t* p, const Word16& v, int jlo, int jhi) {
    if (jlo == 0 && jhi == 16) {
        *reinterpret_cast<uint4*>(p) = make_uint4(v.w[0], v.w[1], v.w[2], v.w[3]);
    } else {
#pragma unroll
        for (int j = 0; j < 16; j++)
            if (j >= jlo && j < jhi) p[j] = (uint8_t)(v.w[j >> 2] >> (8 * (j & 3)));
    }
}

// ---- difference (writer side) ---------------------------------------------------------------
// One CTA per strip.  The strip is rewritten in place from its end towards its start in pieces
// of kDiffThreads words: every thread loads its word and the 4 bytes before it, the CTA meets at a
// barrier, then stores -- so no thread ever reads a byte that was already differenced.
template <int SPP>
__global__ void __launch_bounds__(kDiffThreads)
    slzw_hdiff_kernel(uint8_t* __restrict__ data, const uint64_t* __restrict__ off,
                      const uint64_t* __restrict__ len, uint64_t n, uint32_t row_bytes) {
    constexpr uint32_t kPiece = kDiffThreads * 16;
    const uint32_t step = kPiece % row_bytes;
    for (uint64_t s = blockIdx.x; s < n; s += gridDim.x) {
        const uint64_t b = off[s];
        const uint64_t L = len ? len[s] : off[s + 1] - b;
        if (L == 0) continue;
        uint8_t* sp = data + b;
        const uint32_t skew = (uint32_t)(reinterpret_cast<uintptr_t>(sp) & 15u);
        uint8_t* base = sp - skew;
        const uint64_t nwords = (skew + L + 15) >> 4;
        const uint64_t pieces = (nwords + kDiffThreads - 1) / kDiffThreads;
        // column (position in its row) of the first byte of this thread's word in the last piece
        uint64_t wi = (pieces - 1) * kDiffThreads + threadIdx.x;
        uint32_t col = wi * 16 >= skew ? (uint32_t)((wi * 16 - skew) % row_bytes) : 0;
        for (uint64_t p = pieces; p-- > 0; wi -= kDiffThreads) {
            const bool have = wi < nwords;
            Word16 v, d;
            int jlo = 0, jhi = 0;
            if (have) {
                const uint64_t w16 = wi * 16;
                jlo = w16 >= skew ? 0 : (int)(skew - w16);
                jhi = w16 + 16 <= skew + L ? 16 : (int)(skew + L - w16);
                v = load_word(base + w16, jlo, jhi);
                // the 4 bytes before the word, as far as they belong to the strip
                uint32_t prev = 0;
                if (w16 >= skew + 4) {
                    prev = *reinterpret_cast<const uint32_t*>(base + w16 - 4);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (w16 + j >= skew + 4) prev |= (uint32_t)base[w16 - 4 + j] << (8 * j);
                }
                // bytes at columns < SPP start a row and stay as they are
                uint32_t keep = 0;
                uint32_t j0;  // first row start at or after byte 0 of the word
                if (jlo > 0) {
                    j0 = jlo;
                } else {
                    j0 = col == 0 ? 0 : row_bytes - col;
                    if (col < SPP) keep = (1u << (SPP - col)) - 1;
                }
                for (; j0 < 16; j0 += row_bytes) keep |= ((1u << SPP) - 1) << j0;
                const uint32_t x[5] = {prev, v.w[0], v.w[1], v.w[2], v.w[3]};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t left = SPP == 4 ? x[k] : __funnelshift_l(x[k], x[k + 1], 8 * SPP);
                    const uint32_t m = ((((keep >> (4 * k)) & 0xFu) * 0x00204081u) & 0x01010101u) * 0xFFu;
                    d.w[k] = (__vsub4(v.w[k], left) & ~m) | (v.w[k] & m);
                }
            }
            __syncthreads();
            if (have) store_word(base + wi * 16, d, jlo, jhi);
            col = col >= step ? col - step : col + row_bytes - step;
        }
    }
}

// ---- accumulate (reader side) ---------------------------------------------------------------
// Every row is an independent running sum per channel (mod 256).  The rows of a strip lie back to
// back, so a warp takes a contiguous group of rows and walks it 64 words per pass as one segmented
// scan: a lane sums its 32 bytes per channel (restarting where a row starts inside its word), the
// warp scans the packed per-channel totals, every lane adds what came before it in its own row to
// the bytes ahead of its first row start.  Lane efficiency therefore does
// not depend on the row length.  Channels are counted from byte 0 of the lane's word ("relative");
// 32 and 1024 are multiples of 1, 2 and 4, so for those the relative channel is the same in every
// lane and pass.  For 3 samples per pixel the totals are rotated to row channels and back.
template <int SPP, int K>
__device__ __forceinline__ constexpr uint32_t pattern_selector() {
    uint32_t s = 0;
    for (int i = 0; i < 4; i++) s |= (uint32_t)((4 * K + i) % SPP) << (4 * i);
    return s;
}

__device__ __forceinline__ uint32_t nibble_to_bytes(uint32_t nib) {
    return (((nib & 0xFu) * 0x00204081u) & 0x01010101u) * 0xFFu;
}

// p is the start of a row; len bytes (whole rows, the last one may be short).  A lane takes W
// consecutive words per pass.
template <int SPP, int W>
__device__ __forceinline__ void accumulate_rows(uint8_t* p, uint64_t len, uint32_t row_bytes, int lane) {
    static_assert(W == 1 || W == 2, "row starts of a lane's bytes are kept in one 32-bit mask");
    constexpr int kLaneBytes = 16 * W;
    constexpr uint32_t kPassBytes = kWarpSize * kLaneBytes;
    constexpr uint32_t kFields = 0x00FF00FFu;
    const uint32_t skew = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 15u);
    uint8_t* base = p - skew;
    const uint64_t nwords = (skew + len + 15) >> 4;
    const uint32_t step = kPassBytes % row_bytes;
    // column of byte 0 of this lane's words (bytes before the first row count backwards)
    uint32_t col = (uint32_t)((lane * kLaneBytes + (uint64_t)row_bytes * 16 - skew) % row_bytes);
    uint32_t phase = (lane * kLaneBytes + 48 - skew) % 3;  // row channel of that byte (SPP == 3)
    // per-channel sums of the current row before this pass: channels 0, 2 / 1, 3 in 16-bit fields
    uint32_t carry_lo = 0, carry_hi = 0;

    struct Chunk {
        Word16 v[W];
        int jlo[W], jhi[W];
    };
    auto fetch = [&](uint64_t first_word) {
        Chunk c;
#pragma unroll
        for (int k = 0; k < W; k++) {
            const uint64_t wi = first_word + k;
            c.v[k] = Word16{{0, 0, 0, 0}};
            c.jlo[k] = c.jhi[k] = 0;
            if (wi < nwords) {
                const uint64_t w16 = wi * 16;
                c.jlo[k] = w16 >= skew ? 0 : (int)(skew - w16);
                c.jhi[k] = w16 + 16 <= skew + len ? 16 : (int)(skew + len - w16);
                c.v[k] = load_word(base + w16, c.jlo[k], c.jhi[k]);
            }
        }
        return c;
    };

    Chunk cur = fetch((uint64_t)lane * W);
    for (uint64_t w0 = 0; w0 < nwords; w0 += kWarpSize * W) {
        const uint64_t wi = w0 + (uint64_t)lane * W;
        const Chunk nxt = fetch(wi + kWarpSize * W);  // next pass, in flight during this one
        // row starts inside this lane's bytes, one bit per byte
        uint32_t starts = 0;
        for (uint32_t j0 = col == 0 ? 0 : row_bytes - col; j0 < kLaneBytes; j0 += row_bytes) starts |= 1u << j0;
        uint32_t r[SPP];
#pragma unroll
        for (int c = 0; c < SPP; c++) r[c] = 0;
        Word16 o[W];
#pragma unroll
        for (int k = 0; k < W; k++) o[k] = Word16{{0, 0, 0, 0}};
        if (starts == 0) {
#pragma unroll
            for (int j = 0; j < kLaneBytes; j++) {
                r[j % SPP] += (cur.v[j >> 4].w[(j >> 2) & 3] >> (8 * (j & 3))) & 0xFFu;
                o[j >> 4].w[(j >> 2) & 3] |= (r[j % SPP] & 0xFFu) << (8 * (j & 3));
            }
        } else {
#pragma unroll
            for (int j = 0; j < kLaneBytes; j++) {
                if ((starts >> j) & 1u) {
#pragma unroll
                    for (int c = 0; c < SPP; c++) r[c] = 0;
                }
                r[j % SPP] += (cur.v[j >> 4].w[(j >> 2) & 3] >> (8 * (j & 3))) & 0xFFu;
                o[j >> 4].w[(j >> 2) & 3] |= (r[j % SPP] & 0xFFu) << (8 * (j & 3));
            }
        }
        uint32_t tot = 0;
#pragma unroll
        for (int c = 0; c < SPP; c++) tot |= (r[c] & 0xFFu) << (8 * c);
        uint32_t to_rel = 0x3210;
        if (SPP == 3) {
            // relative channel c is row channel (phase + c) % 3
            const uint32_t to_row = phase == 0 ? 0x3210u : phase == 1 ? 0x3102u : 0x3021u;
            to_rel = phase == 0 ? 0x3210u : phase == 1 ? 0x3021u : 0x3102u;
            tot = __byte_perm(tot, 0, to_row);
        }
        // Segmented exclusive scan.  Sums are mod 256, so a plain scan is enough: what a lane has
        // before it in its own row is (everything before it) - (everything before the last lane
        // below it in which a row started).  The channel sums sit in 16-bit fields, two per
        // register, so that a plain add never carries from one channel into the next (32 lanes of
        // at most 255 each); they are cut back to 8 bits when they are used.
        const uint32_t tot_lo = tot & kFields, tot_hi = (tot >> 8) & kFields;
        uint32_t lo = tot_lo, hi = tot_hi;
#pragma unroll
        for (int dlt = 1; dlt < kWarpSize; dlt <<= 1) {
            const uint32_t t = __shfl_up_sync(kFullMask, lo, dlt);
            const uint32_t u = SPP > 1 ? __shfl_up_sync(kFullMask, hi, dlt) : 0;
            if (lane >= dlt) lo += t, hi += u;
        }
        const uint32_t excl_lo = lo - tot_lo, excl_hi = hi - tot_hi;
        const uint32_t started = __ballot_sync(kFullMask, starts != 0);
        const uint32_t below = started & ((1u << lane) - 1);
        const int seg = below ? 31 - __clz(below) : 0;
        const int last = started ? 31 - __clz(started) : 0;
        const uint32_t cut_lo = __shfl_sync(kFullMask, excl_lo, seg);
        const uint32_t cut_hi = SPP > 1 ? __shfl_sync(kFullMask, excl_hi, seg) : 0;
        const uint32_t b_lo = below ? excl_lo - cut_lo : excl_lo + carry_lo;
        const uint32_t b_hi = below ? excl_hi - cut_hi : excl_hi + carry_hi;
        uint32_t before = (b_lo & kFields) | ((b_hi & kFields) << 8);
        const uint32_t all_lo = __shfl_sync(kFullMask, lo, kWarpSize - 1);
        const uint32_t all_hi = SPP > 1 ? __shfl_sync(kFullMask, hi, kWarpSize - 1) : 0;
        const uint32_t end_lo = __shfl_sync(kFullMask, excl_lo, last);
        const uint32_t end_hi = SPP > 1 ? __shfl_sync(kFullMask, excl_hi, last) : 0;
        carry_lo = (started ? all_lo - end_lo : carry_lo + all_lo) & kFields;
        carry_hi = (started ? all_hi - end_hi : carry_hi + all_hi) & kFields;
        if (SPP == 3) before = __byte_perm(before, 0, to_rel);
        if (wi < nwords) {
            // the bytes ahead of the first row start continue the previous row, the others do not
            const uint32_t ahead = starts ? (1u << (__ffs(starts) - 1)) - 1 : 0xFFFFFFFFu;
#pragma unroll
            for (int k = 0; k < W; k++) {
                uint32_t add[4] = {__byte_perm(before, 0, pattern_selector<SPP, 0>()),
                                   __byte_perm(before, 0, pattern_selector<SPP, 1>()),
                                   __byte_perm(before, 0, pattern_selector<SPP, 2>()),
                                   __byte_perm(before, 0, pattern_selector<SPP, 3>())};
                if (k == 1) {
                    add[0] = __byte_perm(before, 0, pattern_selector<SPP, 4>());
                    add[1] = __byte_perm(before, 0, pattern_selector<SPP, 5>());
                    add[2] = __byte_perm(before, 0, pattern_selector<SPP, 6>());
                    add[3] = __byte_perm(before, 0, pattern_selector<SPP, 7>());
                }
                if (starts) {
#pragma unroll
                    for (int q = 0; q < 4; q++) add[q] &= nibble_to_bytes(ahead >> (16 * k + 4 * q));
                }
#pragma unroll
                for (int q = 0; q < 4; q++) o[k].w[q] = __vadd4(o[k].w[q], add[q]);
                if (wi + k < nwords) store_word(base + (wi + k) * 16, o[k], cur.jlo[k], cur.jhi[k]);
            }
        }
        cur = nxt;
        col += step;
        if (col >= row_bytes) col -= row_bytes;
        phase = (phase + kPassBytes % 3) % 3;
    }
}

template <int SPP>
__global__ void __launch_bounds__(kAccWarps* kWarpSize)
    slzw_hacc_kernel(uint8_t* __restrict__ data, const uint64_t* __restrict__ off,
                     const uint64_t* __restrict__ len, uint64_t n, uint32_t row_bytes) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint64_t s = blockIdx.x; s < n; s += gridDim.x) {
        const uint64_t b = off[s];
        const uint64_t L = len ? len[s] : off[s + 1] - b;
        const uint64_t rows = (L + row_bytes - 1) / row_bytes;
        // the strip's rows in kAccWarps contiguous groups
        const uint64_t r0 = rows * warp / kAccWarps, r1 = rows * (warp + 1) / kAccWarps;
        if (r1 > r0) {
            const uint64_t a = r0 * row_bytes;
            const uint64_t e = r1 * row_bytes < L ? r1 * row_bytes : L;
            accumulate_rows<SPP, kAccLaneWords>(data + b + a, e - a, row_bytes, lane);
        }
    }
}

}  // namespace

// direction 0: difference, 1: accumulate.  samples_per_pixel must be 1..4 and divide row_bytes
// (checked by the caller).
cudaError_t predictor_launch(int direction, uint8_t* data, const uint64_t* off, const uint64_t* len,
                             uint64_t n, uint32_t row_bytes, uint32_t spp, int num_sms,
                             cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    if (direction == 0) {
        const uint64_t cap = (uint64_t)num_sms * 8;
        const unsigned grid = (unsigned)(n < cap ? n : cap);
        switch (spp) {
            case 1: slzw_hdiff_kernel<1><<<grid, kDiffThreads, 0, stream>>>(data, off, len, n, row_bytes); break;
            case 2: slzw_hdiff_kernel<2><<<grid, kDiffThreads, 0, stream>>>(data, off, len, n, row_bytes); break;
            case 3: slzw_hdiff_kernel<3><<<grid, kDiffThreads, 0, stream>>>(data, off, len, n, row_bytes); break;
            default: slzw_hdiff_kernel<4><<<grid, kDiffThreads, 0, stream>>>(data, off, len, n, row_bytes); break;
        }
    } else {
        const uint64_t cap = (uint64_t)num_sms * 16;
        const unsigned grid = (unsigned)(n < cap ? n : cap);
        constexpr int T = kAccWarps * kWarpSize;
        switch (spp) {
            case 1: slzw_hacc_kernel<1><<<grid, T, 0, stream>>>(data, off, len, n, row_bytes); break;
            case 2: slzw_hacc_kernel<2><<<grid, T, 0, stream>>>(data, off, len, n, row_bytes); break;
            case 3: slzw_hacc_kernel<3><<<grid, T, 0, stream>>>(data, off, len, n, row_bytes); break;
            default: slzw_hacc_kernel<4><<<grid, T, 0, stream>>>(data, off, len, n, row_bytes); break;
        }
    }
    return cudaGetLastError();
}

}  // namespace slzw
