// sched_kernels.cu -- stream scheduler and output compaction for sm_100a.
//
// Scheduler: streams are independent and their cost is proportional to their
// length, so the persistent codec kernels consume them largest-first from a work queue
// (longest-processing-time order keeps the tail of the batch short).  The order is produced on
// the device by a counting sort over 1/32-octave size classes: histogram -> exclusive scan of
// the classes from large to small -> scatter.
//
// Compaction: dst_off = exclusive prefix sum of the per-stream sizes (optionally rounded up to
// an alignment), then a warp-per-stream gather that writes aligned 32-bit words whatever the
// relative alignment of source slot and destination (funnel shift of two aligned loads).
#include "slzw_device.cuh"

namespace slzw {

namespace {

constexpr int kSizeClasses = 64 * 32;  // 64 octaves x 32 sub-classes

__device__ __forceinline__ uint32_t size_class(uint64_t len) {
    if (len < 32) return (uint32_t)len;  // octaves 0..4 collapse onto exact sizes
    const int msb = 63 - __clzll((long long)len);
    const uint32_t frac = (uint32_t)((len >> (msb - 5)) & 31u);
    return (uint32_t)msb * 32u + frac;
}

__global__ void sched_histogram_kernel(const uint64_t* __restrict__ off, uint64_t n,
                                       uint32_t* __restrict__ hist) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        atomicAdd(&hist[size_class(off[i + 1] - off[i])], 1u);
}

// one block: hist[c] <- number of streams in classes larger than c (descending exclusive scan)
__global__ void sched_scan_kernel(uint32_t* __restrict__ hist) {
    __shared__ uint32_t part[1024];
    const int t = threadIdx.x;  // 1024 threads, 2 classes each, processed from the top class
    const int c0 = kSizeClasses - 1 - 2 * t, c1 = c0 - 1;
    const uint32_t h0 = hist[c0], h1 = hist[c1];
    part[t] = h0 + h1;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        const uint32_t v = t >= d ? part[t - d] : 0u;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    const uint32_t excl = part[t] - (h0 + h1);
    hist[c0] = excl;
    hist[c1] = excl + h0;
}

__global__ void sched_scatter_kernel(const uint64_t* __restrict__ off, uint64_t n,
                                     uint32_t* __restrict__ cursor, uint32_t* __restrict__ order) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t slot = atomicAdd(&cursor[size_class(off[i + 1] - off[i])], 1u);
        order[slot] = (uint32_t)i;
    }
}

// ---- compaction -----------------------------------------------------------------------------
// single block, chunked scan with a running carry; dst_off has n + 1 entries
__global__ void compact_scan_kernel(const uint64_t* __restrict__ len, uint64_t n, uint64_t align,
                                    uint64_t* __restrict__ dst_off) {
    __shared__ uint64_t part[1024];
    __shared__ uint64_t carry_s;
    const int t = threadIdx.x;
    if (t == 0) carry_s = 0;
    __syncthreads();
    for (uint64_t base = 0; base < n; base += 1024) {
        const uint64_t i = base + t;
        uint64_t v = i < n ? len[i] : 0;
        v = (v + align - 1) / align * align;
        part[t] = v;
        __syncthreads();
        for (int d = 1; d < 1024; d <<= 1) {
            const uint64_t u = t >= d ? part[t - d] : 0;
            __syncthreads();
            part[t] += u;
            __syncthreads();
        }
        const uint64_t carry = carry_s;
        if (i < n) dst_off[i] = carry + part[t] - v;
        __syncthreads();
        if (t == 1023) carry_s = carry + part[1023];
        __syncthreads();
    }
    if (t == 0) dst_off[n] = carry_s;
}

// Copies len bytes src -> dst with aligned 32-bit stores on the destination side.
__device__ __forceinline__ void warp_copy(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src,
                                          uint64_t len, int lane) {
    // head: bring dst to a 4-byte boundary
    uint64_t head = (4 - (reinterpret_cast<uintptr_t>(dst) & 3u)) & 3u;
    if (head > len) head = len;
    if ((uint64_t)lane < head) dst[lane] = src[lane];
    dst += head;
    src += head;
    len -= head;
    const uint64_t nwords = len >> 2;
    const uint32_t sm = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u);
    const uint32_t* s4 = reinterpret_cast<const uint32_t*>(src - sm);
    uint32_t* d4 = reinterpret_cast<uint32_t*>(dst);
    if (sm == 0) {
        for (uint64_t w = lane; w < nwords; w += kWarpSize) d4[w] = s4[w];
    } else {
        const uint32_t sh = sm * 8;
        for (uint64_t w = lane; w < nwords; w += kWarpSize) {
            // s4[w + 1] holds at least one byte of this word, so the load stays inside the source
            const uint32_t lo = s4[w], hi = s4[w + 1];
            d4[w] = __funnelshift_r(lo, hi, sh);
        }
    }
    const uint64_t tail0 = nwords << 2;
    if (tail0 + lane < len) dst[tail0 + lane] = src[tail0 + lane];
}

// One block per stream at a time; the warps of the block take 4 KB pieces of it, so a batch of a
// few long streams (1 MiB GIF frames) spreads over the whole device as well as one of many strips.
constexpr uint64_t kGatherPiece = 4096;

__global__ void compact_gather_kernel(const uint8_t* __restrict__ src,
                                      const uint64_t* __restrict__ src_off,
                                      const uint64_t* __restrict__ len, uint64_t n,
                                      uint8_t* __restrict__ dst, const uint64_t* __restrict__ dst_off) {
    const int lane = threadIdx.x % kWarpSize;
    const uint64_t warp = threadIdx.x / kWarpSize;
    const uint64_t warps = blockDim.x / kWarpSize;
    for (uint64_t i = blockIdx.x; i < n; i += gridDim.x) {
        const uint64_t l = len[i];
        const uint8_t* s = src + src_off[i];
        uint8_t* d = dst + dst_off[i];
        for (uint64_t o = warp * kGatherPiece; o < l; o += warps * kGatherPiece) {
            const uint64_t m = l - o < kGatherPiece ? l - o : kGatherPiece;
            warp_copy(d + o, s + o, m, lane);
        }
    }
}

}  // namespace

int sched_size_classes() { return kSizeClasses; }

// hist: kSizeClasses u32 (scratch), order: n u32
cudaError_t sched_build_order(const uint64_t* off, uint64_t n, uint32_t* hist, uint32_t* order,
                              int num_sms, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(hist, 0, sizeof(uint32_t) * kSizeClasses, stream);
    if (e != cudaSuccess) return e;
    const int threads = 256;
    uint64_t blocks64 = (n + threads - 1) / threads;
    const int blocks = (int)(blocks64 < (uint64_t)(num_sms * 8) ? blocks64 : (uint64_t)(num_sms * 8));
    sched_histogram_kernel<<<blocks, threads, 0, stream>>>(off, n, hist);
    sched_scan_kernel<<<1, 1024, 0, stream>>>(hist);
    sched_scatter_kernel<<<blocks, threads, 0, stream>>>(off, n, hist, order);
    return cudaGetLastError();
}

cudaError_t compact_launch(const uint8_t* src, const uint64_t* src_off, const uint64_t* len,
                           uint64_t n, uint64_t align, uint8_t* dst, uint64_t* dst_off, int num_sms,
                           cudaStream_t stream) {
    compact_scan_kernel<<<1, 1024, 0, stream>>>(len, n, align ? align : 1, dst_off);
    const int threads = 256;
    const uint64_t cap = (uint64_t)num_sms * 8;
    compact_gather_kernel<<<(int)(n < cap ? n : cap), threads, 0, stream>>>(src, src_off, len, n, dst,
                                                                            dst_off);
    return cudaGetLastError();
}

}  // namespace slzw
