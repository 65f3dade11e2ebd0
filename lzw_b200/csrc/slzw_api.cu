// slzw_api.cu -- the C ABI of include/slzw.h: contexts, scheduling, launches, host staging.
//
// There is no CPU code path for the codec in this file: every data-moving entry point ends in
// the sm_100a kernels of encode_kernels.cu / decode_kernels.cu / sched_kernels.cu and returns
// SLZW_RC_NO_DEVICE / SLZW_RC_CUDA when that is impossible.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <new>
#include <chrono>
#include <vector>

#include "slzw_device.cuh"

namespace slzw {
// encode_kernels.cu
void encode_select_config(int c);
cudaError_t encode_configure();
cudaError_t encode_launch(const DevBatch& a, int num_sms, cudaStream_t stream);
size_t encode_table_bytes(uint64_t n, int num_sms);
// decode_kernels.cu
size_t decode_exact_smem_bytes();
int decode_exact_warps_per_cta();
cudaError_t decode_exact_configure();
cudaError_t decode_exact_launch(const DevBatch& a, int grid, cudaStream_t stream);
void decode_select_config(int c);
size_t decode_fast_table_bytes(int num_sms);
cudaError_t decode_fast_configure();
cudaError_t decode_fast_launch(const DevBatch& a, int num_sms, cudaStream_t stream);
// sched_kernels.cu
int sched_size_classes();
cudaError_t sched_build_order(const uint64_t* off, uint64_t n, uint32_t* hist, uint32_t* order,
                              int num_sms, cudaStream_t stream);
cudaError_t predictor_launch(int direction, uint8_t* data, const uint64_t* off, const uint64_t* len,
                             uint64_t n, uint32_t row_bytes, uint32_t spp, int num_sms,
                             cudaStream_t stream);
cudaError_t compact_launch(const uint8_t* src, const uint64_t* src_off, const uint64_t* len,
                           uint64_t n, uint64_t align, uint8_t* dst, uint64_t* dst_off, int num_sms,
                           cudaStream_t stream);
}  // namespace slzw

using namespace slzw;

namespace {

// A grow-only device allocation.
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            e = cudaMalloc(&p, bytes);
            want = bytes;
        }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// Scheduler scratch for one in-flight batch.  Reuse on another CUDA stream waits for `done`.
struct Workspace {
    DevBuf queue;  // kQueueWords x unsigned long long: [0] work queue, [1] retry queue, [2] retry count
    DevBuf hist;   // size classes
    DevBuf order;  // n u32
    DevBuf retry;  // n u32: streams the fast decoder deferred
    DevBuf tables; // fast decoder: dictionaries of the warps that have none in shared memory
    DevBuf enc_tables; // encoder: dictionaries of the lanes in global memory
    cudaEvent_t done = nullptr;
    bool used = false;
};

constexpr int kWorkspaces = 4;
constexpr size_t kQueueWords = 8;  // [4..7]: input bytes encoded per kind of warp (diagnostics)
constexpr int kPipe = 4;
// host-path chunks: at least this many input/output bytes each (a chunk must amortise the tail
// of its longest stream), at most kMaxChunks per call
// measured on config 3 (profiles/r01_e2e_notes.md): the encoder wants larger chunks (every chunk
// pays the tail of its longest streams at 28 streams per SM), the decoder is copy-bound
constexpr uint64_t kEncChunkBytes = 384ull << 20;
constexpr uint64_t kDecChunkBytes = 128ull << 20;
constexpr uint64_t kMaxChunks = 32;

// A grow-only pinned host allocation (small per-chunk result arrays).
struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaHostAlloc(&p, bytes + bytes / 8 + 256, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = bytes + bytes / 8 + 256;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

struct HostSlot {
    cudaStream_t stream = nullptr;
    DevBuf in, out, in_off, out_off, out_len, status, detail, cs, dense, dense_off;
    PinBuf stage;  // pinned staging of the chunk's small arrays (ChunkStage)
    void release() {
        for (DevBuf* b : {&in, &out, &in_off, &out_off, &out_len, &status, &detail, &cs, &dense, &dense_off})
            b->release();
        stage.release();
        if (stream) cudaStreamDestroy(stream);
        stream = nullptr;
    }
};

// Streams and events of the dense encode pipeline (run_host_encode_dense), created on first use.
struct EncPipe {
    cudaStream_t in = nullptr, small = nullptr, dense = nullptr, cmp[2] = {nullptr, nullptr};
    cudaEvent_t ev_in[kPipe] = {}, ev_enc[kPipe] = {}, ev_cmp[kPipe] = {}, ev_small[kPipe] = {}, ev_out[kPipe] = {};
    bool used[kPipe] = {};
    bool ready = false;
    cudaError_t create() {
        if (ready) return cudaSuccess;
        cudaError_t e;
        for (cudaStream_t* st : {&in, &small, &dense, &cmp[0], &cmp[1]})
            if ((e = cudaStreamCreateWithFlags(st, cudaStreamNonBlocking)) != cudaSuccess) return e;
        for (int i = 0; i < kPipe; i++)
            for (cudaEvent_t* ev : {&ev_in[i], &ev_enc[i], &ev_cmp[i], &ev_small[i], &ev_out[i]})
                if ((e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming)) != cudaSuccess) return e;
        ready = true;
        return cudaSuccess;
    }
    void sync_all() {
        for (cudaStream_t st : {in, small, dense, cmp[0], cmp[1]})
            if (st) cudaStreamSynchronize(st);
    }
    void release() {
        for (cudaStream_t* st : {&in, &small, &dense, &cmp[0], &cmp[1]}) {
            if (*st) cudaStreamDestroy(*st);
            *st = nullptr;
        }
        for (int i = 0; i < kPipe; i++)
            for (cudaEvent_t* ev : {&ev_in[i], &ev_enc[i], &ev_cmp[i], &ev_small[i], &ev_out[i]}) {
                if (*ev) cudaEventDestroy(*ev);
                *ev = nullptr;
            }
        ready = false;
    }
};

}  // namespace

// ChunkShape::kHill (below): first chunk in MiB, ratio of the second to the first, growth of the
// following ones up to the chunk size, taper after it
struct HillShape {
    double first_mb = 96, second = 2.0, grow = 1.36, taper = 0.62;
};

struct slzw_ctx {
    int device = 0;
    int num_sms = 0;
    Workspace ws[kWorkspaces];
    int ws_next = 0;
    // host-path staging: kPipe slots, each with its own CUDA stream, so that the H2D copy of one
    // chunk of streams, the kernels of the previous chunk and the D2H copy of the one before
    // overlap (PCIe is full duplex)
    HostSlot pipe[kPipe];
    EncPipe enc_pipe;  // streams and events of the dense encode pipeline
    // pinned encoder input read in place by the kernels instead of staged (SLZW_HOST_ZERO_COPY): 0
    // never; 1 (default) the first chunk of a call only -- the one chunk whose staging copy nothing
    // hides; the kernels read host memory at 23 GB/s, the copy engine stages it at 55 GB/s
    // (profiles/r02_e2e_notes.md); 2 every chunk
    int zero_copy_in = 1;
    int enc_shape = 0;  // 0 taper, 1 hill (SLZW_HOST_ENC_HILL="first_mb:second:grow:taper")
    HillShape hill;
    bool chunk_min_streams = true;  // off when SLZW_HOST_CHUNK_BYTES is set (tests force tiny chunks)
    uint64_t enc_chunk_bytes = kEncChunkBytes;
    uint64_t dec_chunk_bytes = kDecChunkBytes;
    // TIFF Predictor = 2 applied by the host entry points (0 = off), slzw_set_tiff_predictor
    uint32_t pred_row_bytes = 0;
    uint32_t pred_spp = 0;
    uint64_t launches = 0;
    int last_decode_ws = -1;  // workspace of the most recent decode call
    int last_encode_ws = -1;  // workspace of the most recent encode call
    char err[256] = {0};
    std::mutex mu;
    // the staging slots of the host entry points belong to one call at a time: a second call on
    // the same context while one is running is refused (SLZW_RC_INVALID), not interleaved
    std::atomic<bool> host_busy{false};
    // two-phase dense encode (slzw_encode_batch_host_dense_begin / _finish): the encoded chunks stay
    // on the device until the caller knows where they go
    DevBuf shard_dense;
    struct DeferredChunk { uint64_t dev_off, bytes; };
    std::vector<DeferredChunk> deferred;
    uint64_t deferred_total = 0;
};

namespace {

int fail_cuda(slzw_ctx* ctx, cudaError_t e, const char* what) {
    if (ctx) snprintf(ctx->err, sizeof ctx->err, "%s: %s", what, cudaGetErrorString(e));
    return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? SLZW_RC_NO_DEVICE
                                                                         : SLZW_RC_CUDA;
}

#define CK(call, what)                                          \
    do {                                                        \
        cudaError_t e__ = (call);                               \
        if (e__ != cudaSuccess) return fail_cuda(ctx, e__, what); \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

// One host-path call per context at a time.
struct HostCallGuard {
    slzw_ctx* ctx;
    bool ok;
    explicit HostCallGuard(slzw_ctx* c) : ctx(c), ok(c && !c->host_busy.exchange(true, std::memory_order_acquire)) {
        if (c && !ok)
            snprintf(c->err, sizeof c->err, "context is in use by another host call (one context per thread)");
    }
    ~HostCallGuard() {
        if (ok) ctx->host_busy.store(false, std::memory_order_release);
    }
};

bool params_ok(const slzw_params* p) {
    return p && (p->flavour == SLZW_FLAVOUR_VARIABLE || p->flavour == SLZW_FLAVOUR_FIXED ||
                 p->flavour == SLZW_FLAVOUR_VARIABLE_LENIENT);
}

// NVTX range around a phase of a call (visible in Nsight Systems timelines; header-only, a no-op
// when no tool is attached).
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// Picks a workspace, makes `stream` wait for its previous user, builds the processing order.
int prepare(slzw_ctx* ctx, const uint64_t* d_in_off, uint64_t n, cudaStream_t stream,
            Workspace** out_ws) {
    Workspace& w = ctx->ws[ctx->ws_next];
    ctx->ws_next = (ctx->ws_next + 1) % kWorkspaces;
    if (!w.done) CK(cudaEventCreateWithFlags(&w.done, cudaEventDisableTiming), "cudaEventCreate");
    if (w.used) CK(cudaStreamWaitEvent(stream, w.done, 0), "cudaStreamWaitEvent");
    CK(w.queue.reserve(sizeof(unsigned long long) * kQueueWords), "cudaMalloc(queue)");
    CK(w.retry.reserve(sizeof(uint32_t) * n), "cudaMalloc(retry)");
    CK(w.hist.reserve(sizeof(uint32_t) * sched_size_classes()), "cudaMalloc(hist)");
    CK(w.order.reserve(sizeof(uint32_t) * n), "cudaMalloc(order)");
    CK(cudaMemsetAsync(w.queue.p, 0, sizeof(unsigned long long) * kQueueWords, stream),
       "cudaMemsetAsync(queue)");
    CK(sched_build_order(d_in_off, n, (uint32_t*)w.hist.p, (uint32_t*)w.order.p, ctx->num_sms,
                         stream),
       "scheduler launch");
    ctx->launches += 3;
    *out_ws = &w;
    return SLZW_RC_OK;
}

int finish(slzw_ctx* ctx, Workspace* w, cudaStream_t stream) {
    CK(cudaEventRecord(w->done, stream), "cudaEventRecord");
    w->used = true;
    return SLZW_RC_OK;
}

DevBatch make_dev_batch(const slzw_params* params, const slzw_batch* b, const Workspace* w) {
    DevBatch a;
    a.in = b->in;
    a.in_off = b->in_off;
    a.out = b->out;
    a.out_off = b->out_off;
    a.out_len = b->out_len;
    a.status = b->status;
    a.detail = b->detail;
    a.code_size = b->code_size;
    a.n = b->n;
    a.order = (const uint32_t*)w->order.p;
    a.queue = (unsigned long long*)w->queue.p;
    a.retry = (uint32_t*)((unsigned long long*)w->queue.p + 2);
    a.retry_ids = (uint32_t*)w->retry.p;
    a.n_dev = nullptr;
    a.dec_tables = (uint32_t*)w->tables.p;
    a.enc_tables = nullptr;
    a.p = *params;
    return a;
}

int grid_for(const slzw_ctx* ctx, uint64_t n, int warps_per_cta) {
    const uint64_t ctas = (n + warps_per_cta - 1) / warps_per_cta;
    return (int)(ctas < (uint64_t)ctx->num_sms ? ctas : (uint64_t)ctx->num_sms);
}

enum class Op { Encode, Decode, DecodedSizes };

int run_device(slzw_ctx* ctx, const slzw_params* params, const slzw_batch* b, cudaStream_t stream,
               Op op) {
    if (!ctx) return SLZW_RC_INVALID;
    NvtxRange range(op == Op::Encode ? "slzw encode batch (device)"
                                     : op == Op::Decode ? "slzw decode batch (device)" : "slzw decoded sizes (device)");
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (!params_ok(params) || !b) {
        snprintf(ctx->err, sizeof ctx->err, "invalid params or batch");
        return SLZW_RC_INVALID;
    }
    if (b->n == 0) return SLZW_RC_OK;
    const bool needs_out = op != Op::DecodedSizes;
    if (!b->in_off || !b->out_len || !b->status || !b->detail || b->n > 0xFFFFFFFFull ||
        (needs_out && (!b->out || !b->out_off))) {
        snprintf(ctx->err, sizeof ctx->err, "invalid batch (null pointer or n >= 2^32)");
        return SLZW_RC_INVALID;
    }
    DeviceGuard guard(ctx->device);
    if (!guard.ok) return fail_cuda(ctx, cudaGetLastError(), "cudaSetDevice");
    Workspace* w = nullptr;
    int rc = prepare(ctx, b->in_off, b->n, stream, &w);
    if (rc != SLZW_RC_OK) return rc;
    DevBatch a = make_dev_batch(params, b, w);
    if (op == Op::DecodedSizes) {
        a.out = nullptr;
        a.out_off = nullptr;
    }
    if (op == Op::Encode) {
        if (const size_t tb = encode_table_bytes(a.n, ctx->num_sms)) {
            CK(w->enc_tables.reserve(tb + 16384), "cudaMalloc(encode tables)");
            a.enc_tables = (uint8_t*)(((uintptr_t)w->enc_tables.p + 16383) & ~(uintptr_t)16383);
        }
        ctx->last_encode_ws = (int)(w - ctx->ws);
        CK(encode_launch(a, ctx->num_sms, stream), "encode launch");
    } else {
        // fast kernel over the whole batch, then the exact kernel over whatever it deferred
        // (the count lives on the device: the second launch is unconditional and usually idle)
        ctx->last_decode_ws = (int)(w - ctx->ws);
        const bool exact_only = getenv("SLZW_DECODE_EXACT") != nullptr;  // debugging knob
        if (!exact_only) {
            CK(w->tables.reserve(decode_fast_table_bytes(ctx->num_sms)), "cudaMalloc(decode tables)");
            a.dec_tables = (uint32_t*)w->tables.p;
            CK(decode_fast_launch(a, ctx->num_sms, stream), "fast decode launch");
            ctx->launches += 1;
            a.order = a.retry_ids;
            a.n_dev = a.retry;
            a.queue = (unsigned long long*)w->queue.p + 1;
        }
        CK(decode_exact_launch(a, grid_for(ctx, b->n, decode_exact_warps_per_cta()), stream),
           "decode launch");
    }
    ctx->launches += 1;
    return finish(ctx, w, stream);
}

// Pinned (page-locked, mapped) host memory can be read by the kernels in place: returns the device
// alias of `p`, or nullptr for pageable memory.  The encoder reads every input byte exactly once,
// one tile ahead of its use, so its input never needs a staging copy.
const uint8_t* device_alias(const void* p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();  // not an error: plain malloc memory on older drivers
        return nullptr;
    }
    if (attr.type == cudaMemoryTypeHost && attr.devicePointer) return (const uint8_t*)attr.devicePointer;
    return nullptr;
}

// Splits streams [0, n) into chunks of roughly equal weight (weight[i+1] - weight[i] per stream).
// A chunk also has to fill the device (min_streams = streams in flight on it): a chunk with fewer
// streams takes as long as its longest stream whatever its size, so a batch of few long streams
// (config 4: 1 MiB frames, 145 ms each) goes through in few chunks (3.2 -> 11 GB/s end to end).
// shape: how the bytes are spread over the chunks.
//   kEven     equal chunks;
//   kTaper    encode: sizes fall geometrically (x 0.6).  What the call cannot overlap with anything is
//             the copy back of its LAST chunk, and every launch pays the load imbalance of its few
//             streams per warp once: few large chunks first, a small one at the end;
//   kRamp     decode: the first two chunks are a quarter and a half of the others, so that the
//             device-to-host copy (what the decode call is bound by) starts early.
//   kHill     encode: small first chunk (read in place while nothing else could hide its copy),
//             sizes growing as fast as the copy of the next chunk hides behind the kernel of the
//             current one, a plateau of chunk_bytes, then the taper.
enum class ChunkShape { kEven, kTaper, kRamp, kHill };

std::vector<uint64_t> chunk_bounds(const uint64_t* weight, uint64_t n, uint64_t chunk_bytes,
                                   uint64_t min_streams, ChunkShape shape = ChunkShape::kEven,
                                   const HillShape* hill = nullptr) {
    const uint64_t total = weight[n] - weight[0];
    if (shape == ChunkShape::kHill && hill && n >= 2 && total > 0) {
        // sizes in bytes: ramp, plateau, taper; scaled to the total at the end
        const double first = hill->first_mb * 1048576.0, peak = (double)chunk_bytes;
        std::vector<double> up, down;
        for (double c = first; c < peak && up.size() < 12; c *= (up.empty() ? hill->second : hill->grow)) up.push_back(c);
        for (double c = peak * hill->taper; c >= first && down.size() < 12; c *= hill->taper) down.push_back(c);
        auto sum = [](const std::vector<double>& v) { double a = 0; for (double x : v) a += x; return a; };
        // a total too small for the whole hill loses its top
        while (sum(up) + sum(down) > (double)total && up.size() + down.size() > 1) {
            if (!down.empty() && (up.empty() || down.front() >= up.back())) down.erase(down.begin());
            else up.pop_back();
        }
        const double rest = (double)total - sum(up) - sum(down);
        uint64_t plateau = rest > 0 ? (uint64_t)(rest / peak + 0.5) : 0;
        std::vector<double> sizes(up);
        for (uint64_t i = 0; i < plateau && sizes.size() + down.size() < kMaxChunks; i++) sizes.push_back(peak);
        sizes.insert(sizes.end(), down.begin(), down.end());
        const double all = sum(sizes);
        std::vector<uint64_t> cb;
        cb.push_back(0);
        uint64_t i = 0;
        double acc = 0;
        for (size_t c = 0; c + 1 < sizes.size(); c++) {
            acc += sizes[c];
            const uint64_t target = weight[0] + (uint64_t)((double)total * (acc / all));
            while (i < n && weight[i] < target) i++;
            // a chunk that does not fill the device is merged into the next one
            if (i < n && i >= cb.back() + (min_streams ? min_streams / 2 : 1)) cb.push_back(i);
        }
        if (n - cb.back() < (min_streams ? min_streams / 2 : 1) && cb.size() > 1) cb.pop_back();
        cb.push_back(n);
        return cb;
    }
    uint64_t chunks = total / chunk_bytes;
    if (min_streams && chunks > n / min_streams) chunks = n / min_streams;
    if (chunks < 1) chunks = 1;
    if (chunks > kMaxChunks) chunks = kMaxChunks;
    if (chunks > n) chunks = n;
    // cumulative share of the bytes at the end of chunk c
    std::vector<double> upto(chunks, 1.0);
    if (shape == ChunkShape::kTaper && chunks >= 3) {
        if (chunks > 6) chunks = 6;
        upto.assign(chunks, 1.0);
        double w = 1.0, sum = 0.0;
        for (uint64_t c = 0; c < chunks; c++, w *= 0.6) sum += w;
        double acc = 0.0;
        w = 1.0;
        for (uint64_t c = 0; c < chunks; c++, w *= 0.6) {
            acc += w;
            upto[c] = acc / sum;
        }
    } else if (shape == ChunkShape::kRamp && chunks >= 4) {
        const double unit = 1.0 / ((double)chunks - 1.25);  // 0.25 + 0.5 + (chunks - 2) units
        double acc = 0.0;
        for (uint64_t c = 0; c < chunks; c++) {
            acc += (c == 0 ? 0.25 : c == 1 ? 0.5 : 1.0) * unit;
            upto[c] = acc;
        }
    } else {
        for (uint64_t c = 0; c < chunks; c++) upto[c] = (double)(c + 1) / (double)chunks;
    }
    std::vector<uint64_t> cb;
    cb.push_back(0);
    uint64_t i = 0;
    for (uint64_t c = 0; c + 1 < chunks; c++) {
        const uint64_t target = weight[0] + (uint64_t)((double)total * upto[c]);
        while (i < n && weight[i] < target) i++;
        if (i > cb.back() && i < n) cb.push_back(i);
    }
    cb.push_back(n);
    return cb;
}

// Small per-chunk arrays (offsets in, sizes / statuses out) are staged through pinned memory of
// the pipeline slot: a cudaMemcpyAsync from or to pageable memory blocks the host until the
// stream reaches it, which would serialise the pipeline.
struct ChunkStage {
    uint64_t* in_off;    // m + 1
    uint64_t* out_off;   // m + 1
    uint64_t* out_len;   // m   (dense encode: dense_off, m + 1)
    uint32_t* status;    // m
    uint32_t* detail;    // m
    uint8_t* cs;         // m
    static size_t bytes(uint64_t m) { return 8 * (m + 1) * 3 + 4 * m * 2 + m + 64; }
    void bind(void* p, uint64_t m) {
        in_off = (uint64_t*)p;
        out_off = in_off + (m + 1);
        out_len = out_off + (m + 1);
        status = (uint32_t*)(out_len + (m + 1));
        detail = status + m;
        cs = (uint8_t*)(detail + m);
    }
};

// An error in the middle of a pipelined call: copies of earlier chunks into the caller's buffers may
// still be in flight, so the slots are drained before the call returns (the error text is kept).
int drain_pipe(slzw_ctx* ctx, int rc) {
    for (int i = 0; i < kPipe; i++) cudaStreamSynchronize(ctx->pipe[i].stream);
    cudaGetLastError();
    return rc;
}

// Host path: the batch goes through the device in chunks of streams, pipelined over kPipe
// slots (H2D of chunk k+1, kernels of chunk k and D2H of chunk k-1 overlap), results land in the
// caller's buffers.  Pinned host buffers (slzw_host_alloc) make the large copies asynchronous.
int run_host(slzw_ctx* ctx, const slzw_params* params, const slzw_batch* b, Op op) {
    if (!ctx) return SLZW_RC_INVALID;
    if (!params_ok(params) || !b) {
        snprintf(ctx->err, sizeof ctx->err, "invalid params or batch");
        return SLZW_RC_INVALID;
    }
    const uint64_t n = b->n;
    if (n == 0) return SLZW_RC_OK;
    HostCallGuard busy(ctx);
    if (!busy.ok) return SLZW_RC_INVALID;
    const bool needs_out = op != Op::DecodedSizes;
    if (!b->in_off || !b->out_len || !b->status || !b->detail ||
        (needs_out && (!b->out || !b->out_off)) || (b->in_off[n] > 0 && !b->in)) {
        snprintf(ctx->err, sizeof ctx->err, "invalid batch (null pointer)");
        return SLZW_RC_INVALID;
    }
    NvtxRange range(op == Op::Encode ? "slzw encode batch (host pipeline)" : "slzw decode batch (host pipeline)");
    DeviceGuard guard(ctx->device);
    if (!guard.ok) return fail_cuda(ctx, cudaGetLastError(), "cudaSetDevice");
    // chunk by the larger side of the stream (uncompressed bytes)
    const std::vector<uint64_t> cb =
        chunk_bounds(needs_out && op == Op::Decode ? b->out_off : b->in_off, n,
                     op == Op::Encode ? ctx->enc_chunk_bytes : ctx->dec_chunk_bytes,
                     ctx->chunk_min_streams ? (uint64_t)ctx->num_sms * (op == Op::Encode ? 28u : 32u) : 0u,
                     op == Op::Encode ? ChunkShape::kTaper : ChunkShape::kRamp);
    const size_t chunks = cb.size() - 1;
    // encode: pinned input is read in place (the decoder's input is small and its access pattern
    // re-reads tiles, it stays staged)
    // (not with the predictor, which rewrites the device copy of the input)
    const bool predict = ctx->pred_row_bytes != 0 && needs_out;
    const uint8_t* in_alias =
        (op == Op::Encode && ctx->zero_copy_in > 0 && !predict) ? device_alias(b->in) : nullptr;

    const uint8_t* const in_alias_call = in_alias;
    auto enqueue = [&](size_t k) -> int {
        HostSlot& hs = ctx->pipe[k % kPipe];
        cudaStream_t s = hs.stream;
        const uint8_t* const in_alias = (k == 0 || ctx->zero_copy_in >= 2) ? in_alias_call : nullptr;
        const uint64_t s0 = cb[k], s1 = cb[k + 1], m = s1 - s0;
        const uint64_t in_lo = b->in_off[s0], in_hi = b->in_off[s1];
        const uint64_t out_lo = needs_out ? b->out_off[s0] : 0, out_hi = needs_out ? b->out_off[s1] : 0;
        {
            std::lock_guard<std::mutex> lock(ctx->mu);
            if (!in_alias) CK(hs.in.reserve(in_hi - in_lo + 16), "cudaMalloc(in)");
            CK(hs.out.reserve(out_hi - out_lo + 16), "cudaMalloc(out)");
            CK(hs.in_off.reserve(sizeof(uint64_t) * (m + 1)), "cudaMalloc(in_off)");
            CK(hs.out_off.reserve(sizeof(uint64_t) * (m + 1)), "cudaMalloc(out_off)");
            CK(hs.out_len.reserve(sizeof(uint64_t) * m), "cudaMalloc(out_len)");
            CK(hs.status.reserve(sizeof(uint32_t) * m), "cudaMalloc(status)");
            CK(hs.detail.reserve(sizeof(uint32_t) * m), "cudaMalloc(detail)");
            if (b->code_size) CK(hs.cs.reserve(m), "cudaMalloc(code_size)");
            CK(hs.stage.reserve(ChunkStage::bytes(m)), "cudaHostAlloc(stage)");
        }
        ChunkStage st;
        st.bind(hs.stage.p, m);
        memcpy(st.in_off, b->in_off + s0, sizeof(uint64_t) * (m + 1));
        if (needs_out) memcpy(st.out_off, b->out_off + s0, sizeof(uint64_t) * (m + 1));
        if (b->code_size) memcpy(st.cs, b->code_size + s0, m);
        // Offsets stay absolute: the device copies of in/out are biased by -lo instead.
        if (!in_alias && in_hi > in_lo)
            CK(cudaMemcpyAsync(hs.in.p, b->in + in_lo, in_hi - in_lo, cudaMemcpyHostToDevice, s), "H2D in");
        CK(cudaMemcpyAsync(hs.in_off.p, st.in_off, sizeof(uint64_t) * (m + 1), cudaMemcpyHostToDevice, s),
           "H2D in_off");
        if (needs_out)
            CK(cudaMemcpyAsync(hs.out_off.p, st.out_off, sizeof(uint64_t) * (m + 1), cudaMemcpyHostToDevice, s),
               "H2D out_off");
        if (b->code_size) CK(cudaMemcpyAsync(hs.cs.p, st.cs, m, cudaMemcpyHostToDevice, s), "H2D code_size");
        slzw_batch d = {};
        d.in = in_alias ? in_alias : (const uint8_t*)hs.in.p - in_lo;
        d.in_off = (const uint64_t*)hs.in_off.p;
        d.out = needs_out ? (uint8_t*)hs.out.p - out_lo : nullptr;
        d.out_off = needs_out ? (const uint64_t*)hs.out_off.p : nullptr;
        d.out_len = (uint64_t*)hs.out_len.p;
        d.status = (uint32_t*)hs.status.p;
        d.detail = (uint32_t*)hs.detail.p;
        d.code_size = b->code_size ? (const uint8_t*)hs.cs.p : nullptr;
        d.n = m;
        if (predict && op == Op::Encode) {
            CK(predictor_launch(0, (uint8_t*)hs.in.p - in_lo, d.in_off, nullptr, m, ctx->pred_row_bytes,
                                ctx->pred_spp, ctx->num_sms, s), "predictor launch");
            ctx->launches += 1;
        }
        int rc = run_device(ctx, params, &d, s, op);
        if (rc != SLZW_RC_OK) return rc;
        if (predict && op == Op::Decode) {
            CK(predictor_launch(1, d.out, d.out_off, d.out_len, m, ctx->pred_row_bytes, ctx->pred_spp,
                                ctx->num_sms, s), "predictor launch");
            ctx->launches += 1;
        }
        if (needs_out && out_hi > out_lo)
            CK(cudaMemcpyAsync(b->out + out_lo, hs.out.p, out_hi - out_lo, cudaMemcpyDeviceToHost, s), "D2H out");
        CK(cudaMemcpyAsync(st.out_len, hs.out_len.p, sizeof(uint64_t) * m, cudaMemcpyDeviceToHost, s), "D2H out_len");
        CK(cudaMemcpyAsync(st.status, hs.status.p, sizeof(uint32_t) * m, cudaMemcpyDeviceToHost, s), "D2H status");
        CK(cudaMemcpyAsync(st.detail, hs.detail.p, sizeof(uint32_t) * m, cudaMemcpyDeviceToHost, s), "D2H detail");
        return SLZW_RC_OK;
    };
    auto finalize = [&](size_t k) -> int {
        HostSlot& hs = ctx->pipe[k % kPipe];
        CK(cudaStreamSynchronize(hs.stream), "cudaStreamSynchronize");
        const uint64_t s0 = cb[k], m = cb[k + 1] - cb[k];
        ChunkStage st;
        st.bind(hs.stage.p, m);
        memcpy(b->out_len + s0, st.out_len, sizeof(uint64_t) * m);
        memcpy(b->status + s0, st.status, sizeof(uint32_t) * m);
        memcpy(b->detail + s0, st.detail, sizeof(uint32_t) * m);
        return SLZW_RC_OK;
    };

    int rc;
    for (size_t k = 0; k < chunks; k++) {
        if ((rc = enqueue(k)) != SLZW_RC_OK) return drain_pipe(ctx, rc);
        // the slot chunk k+1 will use is the one of chunk k+1-kPipe: finish that chunk now
        if (k + 1 >= (size_t)kPipe && (rc = finalize(k + 1 - kPipe)) != SLZW_RC_OK) return drain_pipe(ctx, rc);
    }
    for (size_t k = chunks >= (size_t)kPipe ? chunks - kPipe + 1 : 0; k < chunks; k++)
        if ((rc = finalize(k)) != SLZW_RC_OK) return drain_pipe(ctx, rc);
    return SLZW_RC_OK;
}

// Host path with dense output: worst-case slots stay on the device, compaction before D2H, so
// only encoded bytes cross the bus.  The host learns a chunk's dense size when its kernels are
// done, places the chunk behind the previous one and starts its copy.
//
// An encode CTA takes its SM whole (all registers, all shared memory), so nothing else runs beside
// it: while an encode kernel that is ready to run is queued anywhere, the compaction of the chunk
// before it does not get an SM until that kernel has drained, and whatever the host issues after
// waiting for the compaction (the next input copy) starts that much later -- measured as 8 ms of
// idle device per two chunks with one stream per chunk (profiles/r02_e2e_notes.md).  Hence:
//   * one stream copies the inputs of up to kPipe chunks ahead, gated by nothing but buffer reuse;
//   * chunks alternate between TWO compute streams, each running encode then compaction: the
//     encode kernel of chunk k+2 is ordered behind the compaction of chunk k, so the compaction
//     gets the SMs that the draining encode kernel of chunk k+1 sets free, and chunk k+2 follows;
//   * the small read-backs (sizes, statuses) and the dense copies have a stream each: a read-back
//     never waits in stream order for the host to place an earlier chunk.
// deferred: the dense chunks stay in ctx->shard_dense (worst-case spacing) and only sizes, statuses
// and details come back; run_host_dense_finish copies them out later.
int run_host_encode_dense(slzw_ctx* ctx, const slzw_params* params, const uint8_t* in,
                          const uint64_t* in_off, uint64_t n, const uint8_t* code_size,
                          uint64_t align, uint8_t* out_dense, uint64_t out_cap, uint64_t* out_off,
                          uint32_t* status, uint32_t* detail, uint64_t* needed, bool deferred = false) {
    if (!ctx) return SLZW_RC_INVALID;
    if (!params_ok(params) || !in_off || !out_off || !status || !detail || (n && !out_dense && !deferred) ||
        (n && in_off[n] > in_off[0] && !in)) {
        snprintf(ctx->err, sizeof ctx->err, "invalid arguments");
        return SLZW_RC_INVALID;
    }
    out_off[0] = 0;
    if (needed) *needed = 0;
    if (deferred) {
        ctx->deferred.clear();
        ctx->deferred_total = 0;
    }
    if (n == 0) return SLZW_RC_OK;
    HostCallGuard busy(ctx);
    if (!busy.ok) return SLZW_RC_INVALID;
    if (align == 0) align = 1;
    NvtxRange range("slzw encode batch, dense (host pipeline)");
    DeviceGuard guard(ctx->device);
    if (!guard.ok) return fail_cuda(ctx, cudaGetLastError(), "cudaSetDevice");
    EncPipe& ep = ctx->enc_pipe;
    {
        std::lock_guard<std::mutex> lock(ctx->mu);
        CK(ep.create(), "cudaStreamCreate / cudaEventCreate (encode pipeline)");
    }
    const std::vector<uint64_t> cb = chunk_bounds(in_off, n, ctx->enc_chunk_bytes,
                                                  ctx->chunk_min_streams ? (uint64_t)ctx->num_sms * 28u : 0u,
                                                  ctx->enc_shape == 1 ? ChunkShape::kHill : ChunkShape::kTaper,
                                                  &ctx->hill);
    const size_t chunks = cb.size() - 1;
    const bool predict = ctx->pred_row_bytes != 0;
    // the first chunk of pinned input may be read in place (nothing could hide its copy), unless
    // the predictor has to rewrite it on the device first
    const uint8_t* const in_alias_call = (ctx->zero_copy_in > 0 && !predict) ? device_alias(in) : nullptr;
    uint64_t hbase = 0;  // dense bytes placed so far
    bool overflow = false;
    // deferred: device offsets of the chunks inside ctx->shard_dense (worst-case spacing)
    std::vector<uint64_t> dev_off(chunks + 1, 0);
    if (deferred) {
        for (size_t k = 0; k < chunks; k++) {
            uint64_t worst = 0;
            for (uint64_t i = cb[k]; i < cb[k + 1]; i++)
                worst += ((slzw_encode_bound(params, in_off[i + 1] - in_off[i]) + 15) & ~15ull) + align;
            dev_off[k + 1] = dev_off[k] + ((worst + 255) & ~255ull);
        }
        std::lock_guard<std::mutex> lock(ctx->mu);
        CK(ctx->shard_dense.reserve(dev_off[chunks] + 256), "cudaMalloc(dense shard)");
    }

    // SLZW_HOST_TRACE=1 (debugging aid): timeline of the call, one line per chunk on stderr.
    // Every event is recorded right behind an operation of its own stream (an event in front of
    // the first copy of a stream queues behind whatever the stream's last engine is doing).
    const bool trace = getenv("SLZW_HOST_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    std::vector<double> t_enq(chunks, 0.0), t_placed(chunks, 0.0);
    const auto t_call = std::chrono::steady_clock::now();
    auto host_ms = [&]() {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call).count();
    };
    if (trace) {
        tev.resize(chunks * 4 + 1);
        for (auto& e : tev) cudaEventCreate(&e);
        cudaEventRecord(tev[chunks * 4], ep.in);
    }
    auto mark = [&](size_t k, int what, cudaStream_t st) {
        if (trace) cudaEventRecord(tev[k * 4 + what], st);
    };
    auto fail = [&](int rc) {
        ep.sync_all();
        for (auto& e : tev) cudaEventDestroy(e);
        return drain_pipe(ctx, rc);
    };
#define CKP(call, what)                                             \
    do {                                                            \
        cudaError_t e_ = (call);                                    \
        if (e_ != cudaSuccess) return fail(fail_cuda(ctx, e_, what)); \
    } while (0)

    auto enqueue = [&](size_t k) -> int {
        const int slot = (int)(k % kPipe);
        HostSlot& hs = ctx->pipe[slot];
        cudaStream_t sc = ep.cmp[k & 1];
        const uint8_t* const in_alias = (k == 0 || ctx->zero_copy_in >= 2) ? in_alias_call : nullptr;
        const uint64_t s0 = cb[k], s1 = cb[k + 1], m = s1 - s0;
        const uint64_t in_lo = in_off[s0], in_hi = in_off[s1];
        {
            std::lock_guard<std::mutex> lock(ctx->mu);
            CK(hs.stage.reserve(ChunkStage::bytes(m)), "cudaHostAlloc(stage)");
        }
        ChunkStage st;
        st.bind(hs.stage.p, m);
        // worst-case slots of the chunk, 16-byte aligned so that the packer's word stores are aligned
        st.out_off[0] = 0;
        for (uint64_t i = 0; i < m; i++) {
            const uint64_t bnd = slzw_encode_bound(params, in_off[s0 + i + 1] - in_off[s0 + i]);
            st.out_off[i + 1] = st.out_off[i] + ((bnd + 15) & ~15ull);
        }
        const uint64_t slot_bytes = st.out_off[m];
        memcpy(st.in_off, in_off + s0, sizeof(uint64_t) * (m + 1));
        if (code_size) memcpy(st.cs, code_size + s0, m);
        {
            std::lock_guard<std::mutex> lock(ctx->mu);
            if (!in_alias) CK(hs.in.reserve(in_hi - in_lo + 16), "cudaMalloc(in)");
            CK(hs.out.reserve(slot_bytes + 16), "cudaMalloc(slots)");
            if (!deferred) CK(hs.dense.reserve(slot_bytes + align * m + 16), "cudaMalloc(dense)");
            CK(hs.in_off.reserve(sizeof(uint64_t) * (m + 1)), "cudaMalloc(in_off)");
            CK(hs.out_off.reserve(sizeof(uint64_t) * (m + 1)), "cudaMalloc(out_off)");
            CK(hs.dense_off.reserve(sizeof(uint64_t) * (m + 1)), "cudaMalloc(dense_off)");
            CK(hs.out_len.reserve(sizeof(uint64_t) * m), "cudaMalloc(out_len)");
            CK(hs.status.reserve(sizeof(uint32_t) * m), "cudaMalloc(status)");
            CK(hs.detail.reserve(sizeof(uint32_t) * m), "cudaMalloc(detail)");
            if (code_size) CK(hs.cs.reserve(m), "cudaMalloc(code_size)");
        }
        if (trace) t_enq[k] = host_ms();
        // input: the slot's previous chunk (k - kPipe) was placed before this call, so its kernels
        // are done with these buffers
        if (!in_alias && in_hi > in_lo)
            CK(cudaMemcpyAsync(hs.in.p, in + in_lo, in_hi - in_lo, cudaMemcpyHostToDevice, ep.in), "H2D in");
        CK(cudaMemcpyAsync(hs.in_off.p, st.in_off, sizeof(uint64_t) * (m + 1), cudaMemcpyHostToDevice, ep.in),
           "H2D in_off");
        CK(cudaMemcpyAsync(hs.out_off.p, st.out_off, sizeof(uint64_t) * (m + 1), cudaMemcpyHostToDevice, ep.in),
           "H2D slots");
        if (code_size) CK(cudaMemcpyAsync(hs.cs.p, st.cs, m, cudaMemcpyHostToDevice, ep.in), "H2D code_size");
        CK(cudaEventRecord(ep.ev_in[slot], ep.in), "cudaEventRecord");
        mark(k, 1, ep.in);
        // kernels
        CK(cudaStreamWaitEvent(sc, ep.ev_in[slot], 0), "cudaStreamWaitEvent");
        // the dense copy of the slot's previous chunk must have left the dense buffer
        if (ep.used[slot] && !deferred) CK(cudaStreamWaitEvent(sc, ep.ev_out[slot], 0), "cudaStreamWaitEvent");
        slzw_batch d = {};
        d.in = in_alias ? in_alias : (const uint8_t*)hs.in.p - in_lo;
        d.in_off = (const uint64_t*)hs.in_off.p;
        d.out = (uint8_t*)hs.out.p;
        d.out_off = (const uint64_t*)hs.out_off.p;
        d.out_len = (uint64_t*)hs.out_len.p;
        d.status = (uint32_t*)hs.status.p;
        d.detail = (uint32_t*)hs.detail.p;
        d.code_size = code_size ? (const uint8_t*)hs.cs.p : nullptr;
        d.n = m;
        if (predict) {
            CK(predictor_launch(0, (uint8_t*)hs.in.p - in_lo, d.in_off, nullptr, m, ctx->pred_row_bytes,
                                ctx->pred_spp, ctx->num_sms, sc), "predictor launch");
            ctx->launches += 1;
        }
        int rc = run_device(ctx, params, &d, sc, Op::Encode);
        if (rc != SLZW_RC_OK) return rc;
        CK(compact_launch(d.out, d.out_off, d.out_len, m, align,
                          deferred ? (uint8_t*)ctx->shard_dense.p + dev_off[k] : (uint8_t*)hs.dense.p,
                          (uint64_t*)hs.dense_off.p, ctx->num_sms, sc), "compaction launch");
        ctx->launches += 2;
        CK(cudaEventRecord(ep.ev_cmp[slot], sc), "cudaEventRecord");
        mark(k, 2, sc);
        // the chunk's dense offsets come back in the out_len area of the stage (m + 1 entries)
        CK(cudaStreamWaitEvent(ep.small, ep.ev_cmp[slot], 0), "cudaStreamWaitEvent");
        CK(cudaMemcpyAsync(st.out_len, hs.dense_off.p, sizeof(uint64_t) * (m + 1), cudaMemcpyDeviceToHost, ep.small),
           "D2H dense_off");
        CK(cudaMemcpyAsync(st.status, hs.status.p, sizeof(uint32_t) * m, cudaMemcpyDeviceToHost, ep.small), "D2H status");
        CK(cudaMemcpyAsync(st.detail, hs.detail.p, sizeof(uint32_t) * m, cudaMemcpyDeviceToHost, ep.small), "D2H detail");
        CK(cudaEventRecord(ep.ev_small[slot], ep.small), "cudaEventRecord");
        ep.used[slot] = true;
        return SLZW_RC_OK;
    };
    // chunk k's kernels are done: place it behind chunk k-1 and start the copy of its bytes
    auto place = [&](size_t k) -> int {
        const int slot = (int)(k % kPipe);
        HostSlot& hs = ctx->pipe[slot];
        CK(cudaEventSynchronize(ep.ev_small[slot]), "cudaEventSynchronize");
        const uint64_t s0 = cb[k], m = cb[k + 1] - cb[k];
        ChunkStage st;
        st.bind(hs.stage.p, m);
        const uint64_t total = st.out_len[m];
        for (uint64_t i = 1; i <= m; i++) out_off[s0 + i] = hbase + st.out_len[i];
        memcpy(status + s0, st.status, sizeof(uint32_t) * m);
        memcpy(detail + s0, st.detail, sizeof(uint32_t) * m);
        if (deferred) {
            ctx->deferred.push_back({dev_off[k], total});
        } else {
            if (hbase + total > out_cap) overflow = true;
            if (!overflow && total)
                CK(cudaMemcpyAsync(out_dense + hbase, hs.dense.p, total, cudaMemcpyDeviceToHost, ep.dense),
                   "D2H dense");
            CK(cudaEventRecord(ep.ev_out[slot], ep.dense), "cudaEventRecord");
        }
        if (trace) t_placed[k] = host_ms();
        mark(k, 3, ep.dense);
        hbase += total;
        return SLZW_RC_OK;
    };

    int rc;
    // up to kPipe chunks in flight; chunk k + kPipe takes the slot of chunk k once that is placed
    for (size_t k = 0; k < chunks && k < (size_t)kPipe; k++)
        if ((rc = enqueue(k)) != SLZW_RC_OK) return fail(rc);
    for (size_t k = 0; k < chunks; k++) {
        if ((rc = place(k)) != SLZW_RC_OK) return fail(rc);
        if (k + kPipe < chunks && (rc = enqueue(k + kPipe)) != SLZW_RC_OK) return fail(rc);
    }
    CKP(cudaStreamSynchronize(ep.dense), "cudaStreamSynchronize");
    CKP(cudaStreamSynchronize(ep.cmp[0]), "cudaStreamSynchronize");
    CKP(cudaStreamSynchronize(ep.cmp[1]), "cudaStreamSynchronize");
#undef CKP
    if (trace) {
        const double t_end = host_ms();
        fprintf(stderr, "slzw trace: encode call %.2f ms on the host clock, %zu chunks\n", t_end, chunks);
        for (size_t k = 0; k < chunks; k++) {
            float t[4] = {0, 0, 0, 0};
            for (int j = 1; j < 4; j++) cudaEventElapsedTime(&t[j], tev[chunks * 4], tev[k * 4 + j]);
            fprintf(stderr, "slzw trace: chunk %zu %7.1f MiB: enqueued %6.2f | input in %6.2f, kernels done %6.2f | placed %6.2f, copied back %6.2f\n",
                    k, (double)(in_off[cb[k + 1]] - in_off[cb[k]]) / 1048576.0, t_enq[k], t[1], t[2],
                    t_placed[k], t[3]);
        }
        for (auto& e : tev) cudaEventDestroy(e);
    }
    if (needed) *needed = hbase;
    if (deferred) ctx->deferred_total = hbase;
    if (overflow) {
        snprintf(ctx->err, sizeof ctx->err, "dense output needs %llu bytes, capacity is %llu",
                 (unsigned long long)hbase, (unsigned long long)out_cap);
        return SLZW_RC_NOMEM;
    }
    return SLZW_RC_OK;
}

// Second phase of the two-phase dense encode: the chunks that run_host_encode_dense(deferred) left
// on the device go to dst back to back, the copies spread over the pipeline's streams.
int run_host_dense_finish(slzw_ctx* ctx, uint8_t* dst, uint64_t cap) {
    if (!ctx) return SLZW_RC_INVALID;
    HostCallGuard busy(ctx);
    if (!busy.ok) return SLZW_RC_INVALID;
    if (ctx->deferred_total > cap || (ctx->deferred_total && !dst)) {
        snprintf(ctx->err, sizeof ctx->err, "dense output needs %llu bytes, capacity is %llu",
                 (unsigned long long)ctx->deferred_total, (unsigned long long)cap);
        return SLZW_RC_NOMEM;
    }
    DeviceGuard guard(ctx->device);
    if (!guard.ok) return fail_cuda(ctx, cudaGetLastError(), "cudaSetDevice");
    uint64_t at = 0;
    for (size_t k = 0; k < ctx->deferred.size(); k++) {
        const auto& c = ctx->deferred[k];
        if (c.bytes)
            CK(cudaMemcpyAsync(dst + at, (const uint8_t*)ctx->shard_dense.p + c.dev_off, c.bytes,
                               cudaMemcpyDeviceToHost, ctx->pipe[k % kPipe].stream),
               "D2H dense shard");
        at += c.bytes;
    }
    for (int i = 0; i < kPipe; i++) CK(cudaStreamSynchronize(ctx->pipe[i].stream), "cudaStreamSynchronize");
    ctx->deferred.clear();
    ctx->deferred_total = 0;
    return SLZW_RC_OK;
}

int run_single(slzw_ctx* ctx, const slzw_params* params, const uint8_t* in, uint64_t n,
               uint8_t* out, uint64_t cap, uint64_t* out_len, uint32_t* detail, Op op) {
    uint64_t in_off[2] = {0, n};
    uint64_t out_off[2] = {0, cap};
    uint64_t len = 0;
    uint32_t st = 0, det = 0;
    uint8_t dummy = 0;
    slzw_batch b = {};
    b.in = in ? in : &dummy;
    b.in_off = in_off;
    b.out = out ? out : &dummy;
    b.out_off = out_off;
    b.out_len = &len;
    b.status = &st;
    b.detail = &det;
    b.n = 1;
    if (!out) out_off[1] = 0;
    int rc = run_host(ctx, params, &b, op);
    if (rc != SLZW_RC_OK) return rc;
    if (out_len) *out_len = len;
    if (detail) *detail = det;
    return (int)st;
}

}  // namespace

extern "C" {

uint32_t slzw_version(void) { return (SLZW_VERSION_MAJOR << 16) | SLZW_VERSION_MINOR; }

int slzw_create(int device, slzw_ctx** out) {
    if (!out) return SLZW_RC_INVALID;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0 || device < 0 || device >= count) return SLZW_RC_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SLZW_RC_NO_DEVICE;
    // the kernels are built for sm_100a only, which has no forward-compatible PTX (sm_101 / sm_103
    // parts cannot run them)
    if (prop.major != 10 || prop.minor != 0) return SLZW_RC_NO_DEVICE;
    slzw_ctx* ctx = new (std::nothrow) slzw_ctx;
    if (!ctx) return SLZW_RC_NOMEM;
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    {
        const char* e = getenv("SLZW_ENC_CONFIG");  // tests / tuning knob, process-wide; unset = by batch size
        encode_select_config(e ? atoi(e) : 0);
    }
    if (const char* e = getenv("SLZW_DEC_CONFIG")) decode_select_config(atoi(e));  // tuning knob
    if (const char* e = getenv("SLZW_HOST_ZERO_COPY")) ctx->zero_copy_in = atoi(e);
    if (const char* e = getenv("SLZW_HOST_ENC_HILL")) {  // tuning knob: "first_mb:second:grow:taper[:peak_mb]"
        double a, b, c, d, pk = 0;
        const int k = sscanf(e, "%lf:%lf:%lf:%lf:%lf", &a, &b, &c, &d, &pk);
        if (k >= 4 && a > 0 && b >= 1 && c > 1 && d > 0 && d < 1) {
            ctx->enc_shape = 1;
            ctx->hill.first_mb = a;
            ctx->hill.second = b;
            ctx->hill.grow = c;
            ctx->hill.taper = d;
            if (k == 5 && pk >= 1) ctx->enc_chunk_bytes = (uint64_t)(pk * 1048576.0);
        }
    }
    if (const char* e = getenv("SLZW_HOST_CHUNK_BYTES")) {
        const long long v = atoll(e);  // tests use tiny chunks
        if (v > 0) {
            ctx->enc_chunk_bytes = ctx->dec_chunk_bytes = (uint64_t)v;
            ctx->chunk_min_streams = false;
        }
    }
    DeviceGuard guard(device);
    if (!guard.ok || encode_configure() != cudaSuccess || decode_exact_configure() != cudaSuccess ||
        decode_fast_configure() != cudaSuccess) {
        delete ctx;
        return SLZW_RC_CUDA;
    }
    for (int i = 0; i < kPipe; i++) {
        if (cudaStreamCreateWithFlags(&ctx->pipe[i].stream, cudaStreamNonBlocking) != cudaSuccess) {
            for (int j = 0; j < kPipe; j++) ctx->pipe[j].release();
            delete ctx;
            return SLZW_RC_CUDA;
        }
    }
    *out = ctx;
    return SLZW_RC_OK;
}

void slzw_destroy(slzw_ctx* ctx) {
    if (!ctx) return;
    {
        DeviceGuard guard(ctx->device);
        cudaDeviceSynchronize();
        for (auto& w : ctx->ws) {
            w.queue.release();
            w.hist.release();
            w.order.release();
            w.retry.release();
            w.tables.release();
            w.enc_tables.release();
            if (w.done) cudaEventDestroy(w.done);
        }
        for (int i = 0; i < kPipe; i++) ctx->pipe[i].release();
        ctx->enc_pipe.release();
        ctx->shard_dense.release();
    }
    delete ctx;
}

const char* slzw_last_error(const slzw_ctx* ctx) { return ctx ? ctx->err : "null context"; }
uint64_t slzw_kernel_launches(const slzw_ctx* ctx) { return ctx ? ctx->launches : 0; }

uint64_t slzw_last_deferred(slzw_ctx* ctx, uint32_t* ids, uint64_t cap) {
    if (!ctx || ctx->last_decode_ws < 0) return 0;
    std::lock_guard<std::mutex> lock(ctx->mu);
    DeviceGuard guard(ctx->device);
    if (!guard.ok || cudaDeviceSynchronize() != cudaSuccess) return 0;
    const Workspace& w = ctx->ws[ctx->last_decode_ws];
    uint32_t count = 0;
    if (cudaMemcpy(&count, (const unsigned long long*)w.queue.p + 2, sizeof count,
                   cudaMemcpyDeviceToHost) != cudaSuccess)
        return 0;
    const uint64_t m = count < cap ? count : cap;
    if (ids && m && cudaMemcpy(ids, w.retry.p, sizeof(uint32_t) * m, cudaMemcpyDeviceToHost) != cudaSuccess)
        return 0;
    return count;
}

int slzw_last_encode_shares(slzw_ctx* ctx, uint64_t bytes[4]) {
    if (!ctx || !bytes || ctx->last_encode_ws < 0) return SLZW_RC_INVALID;
    std::lock_guard<std::mutex> lock(ctx->mu);
    DeviceGuard guard(ctx->device);
    if (!guard.ok) return fail_cuda(ctx, cudaGetLastError(), "cudaSetDevice");
    CK(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
    CK(cudaMemcpy(bytes, (const unsigned long long*)ctx->ws[ctx->last_encode_ws].queue.p + 4,
                  4 * sizeof(uint64_t), cudaMemcpyDeviceToHost),
       "cudaMemcpy(encode shares)");
    return SLZW_RC_OK;
}

int slzw_encode_batch_device(slzw_ctx* ctx, const slzw_params* params, const slzw_batch* batch,
                             void* cuda_stream) {
    return run_device(ctx, params, batch, (cudaStream_t)cuda_stream, Op::Encode);
}

int slzw_decode_batch_device(slzw_ctx* ctx, const slzw_params* params, const slzw_batch* batch,
                             void* cuda_stream) {
    return run_device(ctx, params, batch, (cudaStream_t)cuda_stream, Op::Decode);
}

int slzw_decoded_sizes_batch_device(slzw_ctx* ctx, const slzw_params* params,
                                    const slzw_batch* batch, void* cuda_stream) {
    return run_device(ctx, params, batch, (cudaStream_t)cuda_stream, Op::DecodedSizes);
}

int slzw_encode_batch_host(slzw_ctx* ctx, const slzw_params* params, const slzw_batch* batch) {
    return run_host(ctx, params, batch, Op::Encode);
}

int slzw_decode_batch_host(slzw_ctx* ctx, const slzw_params* params, const slzw_batch* batch) {
    return run_host(ctx, params, batch, Op::Decode);
}

int slzw_decoded_sizes_batch_host(slzw_ctx* ctx, const slzw_params* params, const slzw_batch* batch) {
    return run_host(ctx, params, batch, Op::DecodedSizes);
}

int slzw_encode_batch_host_dense(slzw_ctx* ctx, const slzw_params* params, const uint8_t* in,
                                 const uint64_t* in_off, uint64_t n, const uint8_t* code_size,
                                 uint64_t align, uint8_t* out_dense, uint64_t out_cap,
                                 uint64_t* out_off, uint32_t* status, uint32_t* detail,
                                 uint64_t* needed) {
    return run_host_encode_dense(ctx, params, in, in_off, n, code_size, align, out_dense, out_cap,
                                 out_off, status, detail, needed);
}

int slzw_encode_batch_host_dense_begin(slzw_ctx* ctx, const slzw_params* params, const uint8_t* in,
                                       const uint64_t* in_off, uint64_t n, const uint8_t* code_size,
                                       uint64_t align, uint64_t* out_off, uint32_t* status,
                                       uint32_t* detail, uint64_t* total) {
    return run_host_encode_dense(ctx, params, in, in_off, n, code_size, align, nullptr, ~0ull, out_off,
                                 status, detail, total, true);
}

int slzw_encode_batch_host_dense_finish(slzw_ctx* ctx, uint8_t* out_dense, uint64_t out_cap) {
    return run_host_dense_finish(ctx, out_dense, out_cap);
}

int slzw_encode(slzw_ctx* ctx, const slzw_params* params, const uint8_t* in, uint64_t n,
                uint8_t* out, uint64_t cap, uint64_t* out_len, uint32_t* detail) {
    return run_single(ctx, params, in, n, out, cap, out_len, detail, Op::Encode);
}

int slzw_decode(slzw_ctx* ctx, const slzw_params* params, const uint8_t* in, uint64_t n,
                uint8_t* out, uint64_t cap, uint64_t* out_len, uint32_t* detail) {
    return run_single(ctx, params, in, n, out, cap, out_len, detail,
                      out ? Op::Decode : Op::DecodedSizes);
}

uint64_t slzw_encode_bound(const slzw_params* params, uint64_t n) {
    (void)params;
    // One code per input byte at most (encoder.rs:322-324), plus the leading clear, one clear
    // per dictionary reset (a reset needs >= 3838 emitted codes: 4096 - 258), the final prefix
    // and EOI; every code <= 12 bits; fill() pads to a byte.
    const uint64_t codes = n + 3 + n / 3838 + 1;
    return (codes * 12 + 7) / 8;
}

int slzw_compact_device(slzw_ctx* ctx, const uint8_t* src, const uint64_t* src_off,
                        const uint64_t* len, uint64_t n, uint64_t align, uint8_t* dst,
                        uint64_t* dst_off, void* cuda_stream) {
    if (!ctx) return SLZW_RC_INVALID;
    if (!dst_off || (n && (!src || !src_off || !len || !dst))) {
        snprintf(ctx->err, sizeof ctx->err, "invalid compaction arguments");
        return SLZW_RC_INVALID;
    }
    DeviceGuard guard(ctx->device);
    if (!guard.ok) return fail_cuda(ctx, cudaGetLastError(), "cudaSetDevice");
    CK(compact_launch(src, src_off, len, n, align, dst, dst_off, ctx->num_sms,
                      (cudaStream_t)cuda_stream), "compaction launch");
    ctx->launches += 2;
    return SLZW_RC_OK;
}

static bool predictor_args_ok(uint32_t row_bytes, uint32_t spp) {
    return row_bytes > 0 && spp >= 1 && spp <= 4 && row_bytes % spp == 0;
}

int slzw_tiff_predictor_device(slzw_ctx* ctx, int direction, uint8_t* data, const uint64_t* off,
                               const uint64_t* len, uint64_t n, uint32_t row_bytes,
                               uint32_t samples_per_pixel, void* cuda_stream) {
    if (!ctx) return SLZW_RC_INVALID;
    if ((direction != SLZW_PREDICTOR_DIFFERENCE && direction != SLZW_PREDICTOR_ACCUMULATE) ||
        !predictor_args_ok(row_bytes, samples_per_pixel) || (n && (!data || !off))) {
        snprintf(ctx->err, sizeof ctx->err,
                 "invalid predictor arguments (8-bit samples, 1..4 per pixel, row_bytes a multiple)");
        return SLZW_RC_INVALID;
    }
    DeviceGuard guard(ctx->device);
    if (!guard.ok) return fail_cuda(ctx, cudaGetLastError(), "cudaSetDevice");
    CK(predictor_launch(direction, data, off, len, n, row_bytes, samples_per_pixel, ctx->num_sms,
                        (cudaStream_t)cuda_stream), "predictor launch");
    ctx->launches += n ? 1 : 0;
    return SLZW_RC_OK;
}

int slzw_set_tiff_predictor(slzw_ctx* ctx, uint32_t row_bytes, uint32_t samples_per_pixel) {
    if (!ctx) return SLZW_RC_INVALID;
    if (row_bytes == 0 && samples_per_pixel == 0) {
        ctx->pred_row_bytes = ctx->pred_spp = 0;
        return SLZW_RC_OK;
    }
    if (!predictor_args_ok(row_bytes, samples_per_pixel)) {
        snprintf(ctx->err, sizeof ctx->err,
                 "invalid predictor arguments (8-bit samples, 1..4 per pixel, row_bytes a multiple)");
        return SLZW_RC_INVALID;
    }
    ctx->pred_row_bytes = row_bytes;
    ctx->pred_spp = samples_per_pixel;
    return SLZW_RC_OK;
}

void* slzw_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}

void slzw_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int slzw_status_message(int is_decoder, uint32_t status, uint32_t detail, uint8_t code_size,
                        char* buf, size_t buf_len) {
    if (!buf || buf_len == 0) return 0;
    int w = 0;
    switch (status) {
        case SLZW_OK:
            w = snprintf(buf, buf_len, "ok");
            break;
        case SLZW_ERR_CODE_SIZE:  // encoder.rs:35-37 (trailing period) vs decoder.rs:31-33
            w = snprintf(buf, buf_len, is_decoder ? "Code size must be between 2 and 8, was %u"
                                                  : "Code size must be between 2 and 8, was %u.",
                         detail);
            break;
        case SLZW_ERR_UNEXPECTED_CODE:  // encoder.rs:38-41, decoder.rs:34-36
            if (is_decoder)
                w = snprintf(buf, buf_len, "Unexpected code while decompressing: %u", detail);
            else
                w = snprintf(buf, buf_len,
                             "Unexpected code %u. For code size %u, data should be < %u.", detail,
                             (unsigned)code_size, 1u << code_size);
            break;
        case SLZW_ERR_MISSING_CLEAR_CODE:  // decoder.rs:37-39 (sic)
            w = snprintf(buf, buf_len, "Dictionnary growing past 4096, expected CLEAR_CODE missing");
            break;
        case SLZW_ERR_IO_UNEXPECTED_EOF:  // std::io::ErrorKind::UnexpectedEof via read_exact
            w = snprintf(buf, buf_len, "failed to fill whole buffer");
            break;
        case SLZW_ERR_IO_WRITE_ZERO:  // std::io::ErrorKind::WriteZero via write_all
            w = snprintf(buf, buf_len, "failed to write whole buffer");
            break;
        case SLZW_ERR_REFERENCE_PANIC:
            w = snprintf(buf, buf_len, "the reference implementation panics on this input");
            break;
        default:
            w = snprintf(buf, buf_len, "unknown status %u", status);
    }
    if (w < 0) return 0;
    return (size_t)w < buf_len ? w : (int)buf_len - 1;
}

}  // extern "C"
