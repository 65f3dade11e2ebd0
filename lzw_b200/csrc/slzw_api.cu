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
#include <algorithm>
#include <atomic>
#include <mutex>
#include <new>
#include <chrono>
#include <thread>
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
constexpr uint64_t kEncChunkBytes = 320ull << 20;
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

// Buffers of the streaming dense encode (run_host_encode_stream): the whole call's input, slots and
// dense output on the device, its small arrays, the pinned staging of those, and the window flags
// the kernel raises in mapped host memory.
constexpr int kStreamRing = 8;  // windows whose compaction may be in flight ahead of the one being placed
struct StreamBufs {
    DevBuf in, slots, dense, in_off, out_off, out_len, status, detail, cs, order, win_of, win_count, ctl, dense_off;
    PinBuf meta, flags;
    cudaEvent_t ev_meta = nullptr, ev_cmp[kStreamRing] = {}, ev_small[kStreamRing] = {};
    cudaError_t create_events() {
        if (ev_meta) return cudaSuccess;
        cudaError_t e;
        if ((e = cudaEventCreateWithFlags(&ev_meta, cudaEventDisableTiming)) != cudaSuccess) return e;
        for (int i = 0; i < kStreamRing; i++) {
            if ((e = cudaEventCreateWithFlags(&ev_cmp[i], cudaEventDisableTiming)) != cudaSuccess) return e;
            if ((e = cudaEventCreateWithFlags(&ev_small[i], cudaEventDisableTiming)) != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    void release() {
        for (DevBuf* b : {&in, &slots, &dense, &in_off, &out_off, &out_len, &status, &detail, &cs, &order, &win_of,
                          &win_count, &ctl, &dense_off})
            b->release();
        meta.release();
        flags.release();
        if (ev_meta) cudaEventDestroy(ev_meta);
        ev_meta = nullptr;
        for (int i = 0; i < kStreamRing; i++) {
            if (ev_cmp[i]) cudaEventDestroy(ev_cmp[i]);
            if (ev_small[i]) cudaEventDestroy(ev_small[i]);
            ev_cmp[i] = ev_small[i] = nullptr;
        }
    }
};

}  // namespace

// ChunkShape::kHill (below): first chunk in MiB, ratio of the second to the first, growth of the
// following ones up to the chunk size, taper after it
struct HillShape {
    double first_mb = 80, second = 1.5, grow = 1.5, taper = 0.7;
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
    StreamBufs sb;     // streaming dense encode
    int enc_stream = 1;                       // dense encode of host batches: 1 streaming, 0 chunked (SLZW_HOST_ENC_STREAM)
    uint64_t stream_window_bytes = 32ull << 20;  // SLZW_HOST_WINDOW_MB
    uint64_t stream_max_bytes = 12ull << 30;  // larger calls take the chunked pipeline
    int stream_reserved_sms = 4;              // SMs the streaming encode launch leaves to the compactions (SLZW_HOST_STREAM_SMS)
    // pinned encoder input read in place by the kernels instead of staged (SLZW_HOST_ZERO_COPY): 0
    // never; 1 (default) the first chunk of a call only -- the one chunk whose staging copy nothing
    // hides; the kernels read host memory at 23 GB/s, the copy engine stages it at 55 GB/s
    // (profiles/r02_e2e_notes.md); 2 every chunk
    int zero_copy_in = 1;
    int enc_shape = 1;  // dense encode: 0 taper, 1 hill (SLZW_HOST_ENC_HILL="first_mb:second:grow:taper[:peak_mb]", "taper")
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
    const uint8_t* deferred_base = nullptr;  // device buffer the deferred chunks live in
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
    a.sc = StreamCtl{};
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

// SLZW_HOST_TRACE=1 (debugging aid): timeline of a host-pipeline call, one line per chunk on stderr:
// host clock when the chunk was enqueued / placed, device clock (relative to the call's first
// event) when its input was in, its kernels were done and its output was copied back.  Every event
// is recorded right behind an operation of its own stream (an event in front of the first copy of
// a stream queues behind whatever the stream's last engine is doing and would delay the copy).
struct HostTrace {
    bool on = false;
    size_t chunks = 0;
    std::vector<cudaEvent_t> ev;
    std::vector<double> t_enq, t_placed;
    std::chrono::steady_clock::time_point t0;
    void begin(size_t n_chunks, cudaStream_t first) {
        on = getenv("SLZW_HOST_TRACE") != nullptr;
        t0 = std::chrono::steady_clock::now();
        if (!on) return;
        chunks = n_chunks;
        ev.resize(chunks * 3 + 1);
        for (auto& e : ev) cudaEventCreate(&e);
        t_enq.assign(chunks, 0.0);
        t_placed.assign(chunks, 0.0);
        cudaEventRecord(ev[chunks * 3], first);
    }
    double host_ms() const {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    void enqueued(size_t k) { if (on) t_enq[k] = host_ms(); }
    void placed(size_t k) { if (on) t_placed[k] = host_ms(); }
    void mark(size_t k, int what, cudaStream_t st) { if (on) cudaEventRecord(ev[k * 3 + what], st); }
    void report(const char* call, const uint64_t* weight, const std::vector<uint64_t>& cb) {
        if (!on) return;
        fprintf(stderr, "slzw trace: %s call %.2f ms on the host clock, %zu chunks\n", call, host_ms(), chunks);
        for (size_t k = 0; k < chunks; k++) {
            float t[3] = {0, 0, 0};
            for (int j = 0; j < 3; j++) cudaEventElapsedTime(&t[j], ev[chunks * 3], ev[k * 3 + j]);
            fprintf(stderr, "slzw trace: chunk %zu %7.1f MiB: enqueued %6.2f | input in %6.2f, kernels done %6.2f | placed %6.2f, copied back %6.2f\n",
                    k, (double)(weight[cb[k + 1]] - weight[cb[k]]) / 1048576.0, t_enq[k], t[0], t[1], t_placed[k], t[2]);
        }
    }
    ~HostTrace() {
        for (auto& e : ev) cudaEventDestroy(e);
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
                     op == Op::Encode ? (ctx->enc_shape == 1 ? ChunkShape::kHill : ChunkShape::kTaper)
                                      : ChunkShape::kRamp,
                     &ctx->hill);
    const size_t chunks = cb.size() - 1;
    // encode: pinned input is read in place (the decoder's input is small and its access pattern
    // re-reads tiles, it stays staged)
    // (not with the predictor, which rewrites the device copy of the input)
    const bool predict = ctx->pred_row_bytes != 0 && needs_out;
    const uint8_t* in_alias =
        (op == Op::Encode && ctx->zero_copy_in > 0 && !predict) ? device_alias(b->in) : nullptr;

    HostTrace tr;
    tr.begin(chunks, ctx->pipe[0].stream);
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
        tr.enqueued(k);
        // Offsets stay absolute: the device copies of in/out are biased by -lo instead.
        if (!in_alias && in_hi > in_lo)
            CK(cudaMemcpyAsync(hs.in.p, b->in + in_lo, in_hi - in_lo, cudaMemcpyHostToDevice, s), "H2D in");
        CK(cudaMemcpyAsync(hs.in_off.p, st.in_off, sizeof(uint64_t) * (m + 1), cudaMemcpyHostToDevice, s),
           "H2D in_off");
        if (needs_out)
            CK(cudaMemcpyAsync(hs.out_off.p, st.out_off, sizeof(uint64_t) * (m + 1), cudaMemcpyHostToDevice, s),
               "H2D out_off");
        if (b->code_size) CK(cudaMemcpyAsync(hs.cs.p, st.cs, m, cudaMemcpyHostToDevice, s), "H2D code_size");
        tr.mark(k, 0, s);
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
        tr.mark(k, 1, s);
        if (needs_out && out_hi > out_lo)
            CK(cudaMemcpyAsync(b->out + out_lo, hs.out.p, out_hi - out_lo, cudaMemcpyDeviceToHost, s), "D2H out");
        tr.mark(k, 2, s);
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
        tr.placed(k);
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
    tr.report(op == Op::Encode ? "encode" : "decode", needs_out && op == Op::Decode ? b->out_off : b->in_off, cb);
    return SLZW_RC_OK;
}

constexpr int kStreamFallback = 1;  // run_host_encode_stream: not enough device memory, nothing was started

// Streaming dense encode of a host batch: ONE encode launch for the whole call.
//
// The chunked pipeline below pays for its chunk boundaries: an encode CTA owns its SM until its 28
// warps are done, a warp whose chunk has run dry cannot take a stream of the next chunk, so a chunk
// that gives every warp only a stream or two leaves half of the dictionaries idle (the first and
// the last chunks of a call, which have to be small: 25 ms for the first 0.38 GB of config 3), and
// the compaction of a chunk cannot run before the encode kernel behind it drains.  Here the input
// is cut into windows (16, 32, then 64 MiB) that the copy engine delivers back to back, each
// followed by a 4-byte copy that advances a device word (`avail`); the kernel's warps take streams
// from one queue over all windows -- largest first inside a window -- and wait for `avail` only when
// they have caught up with the copies (StreamCtl, slzw_device.cuh).  The launch leaves
// stream_reserved_sms SMs free; the warp that finishes a window's last stream raises the window's
// flag in mapped host memory, and the host starts that window's compaction (on the free SMs),
// reads its sizes back, places it behind the previous window and copies it out.
int run_host_encode_stream(slzw_ctx* ctx, const slzw_params* params, const uint8_t* in,
                           const uint64_t* in_off, uint64_t n, const uint8_t* code_size, uint64_t align,
                           uint8_t* out_dense, uint64_t out_cap, uint64_t* out_off, uint32_t* status,
                           uint32_t* detail, uint64_t* needed, bool deferred) {
    NvtxRange range("slzw encode batch, dense (streaming host pipeline)");
    EncPipe& ep = ctx->enc_pipe;
    StreamBufs& sb = ctx->sb;
    {
        std::lock_guard<std::mutex> lock(ctx->mu);
        CK(ep.create(), "cudaStreamCreate / cudaEventCreate (encode pipeline)");
        CK(sb.create_events(), "cudaEventCreate (streaming encode)");
    }
    const uint64_t in_lo = in_off[0], total = in_off[n] - in_off[0];
    // ---- windows: stream ranges [wb[w], wb[w + 1]) ----
    std::vector<uint64_t> wb;
    {
        uint64_t target = ctx->chunk_min_streams ? ctx->stream_window_bytes : ctx->enc_chunk_bytes;
        if (target < 1) target = 1;
        while (total / target > 4000) target *= 2;  // window indices are 16 bits
        uint64_t size = ctx->chunk_min_streams && target > (16ull << 20) ? (16ull << 20) : target;
        wb.push_back(0);
        uint64_t i = 0;
        while (i < n) {
            const uint64_t end = in_off[i] + size;
            uint64_t j = i + 1;
            while (j < n && in_off[j + 1] <= end) j++;
            wb.push_back(j);
            i = j;
            size = size * 2 < target ? size * 2 : target;
        }
    }
    const size_t W = wb.size() - 1;
    // ---- buffers ----
    const size_t ctl_bytes = 128 + 4 * W;
    auto up8 = [](size_t v) { return (v + 7) & ~size_t(7); };
    const size_t meta_bytes = up8(8 * (n + 1)) * 2 + up8(4 * n) + up8(2 * n) + up8(4 * W) * 2 + up8(ctl_bytes) + up8(n) +
                              up8(8 * (n + W)) + up8(4 * n) * 2 + 64;
    {
        std::lock_guard<std::mutex> lock(ctx->mu);
        // the three large buffers first; a device that cannot hold them sends the call to the
        // chunked pipeline.  Slots: slzw_encode_bound is below 1.5004 * len + 7, + 15 for alignment
        const uint64_t slots_max = total + total / 2 + total / 2048 + 32 * n + 4096;
        if (sb.in.reserve(total + 16) != cudaSuccess || sb.slots.reserve(slots_max + 16) != cudaSuccess ||
            sb.dense.reserve(slots_max + align * n + 256) != cudaSuccess) {
            cudaGetLastError();
            sb.in.release();
            sb.slots.release();
            sb.dense.release();
            return kStreamFallback;
        }
        CK(sb.meta.reserve(meta_bytes), "cudaHostAlloc(meta)");
        CK(sb.flags.reserve(4 * (W + 1)), "cudaHostAlloc(flags)");
        CK(sb.in_off.reserve(8 * (n + 1)), "cudaMalloc(in_off)");
        CK(sb.out_off.reserve(8 * (n + 1)), "cudaMalloc(out_off)");
        CK(sb.out_len.reserve(8 * n), "cudaMalloc(out_len)");
        CK(sb.status.reserve(4 * n), "cudaMalloc(status)");
        CK(sb.detail.reserve(4 * n), "cudaMalloc(detail)");
        CK(sb.order.reserve(4 * n), "cudaMalloc(order)");
        CK(sb.win_of.reserve(2 * n), "cudaMalloc(win_of)");
        CK(sb.win_count.reserve(4 * W), "cudaMalloc(win_count)");
        CK(sb.ctl.reserve(ctl_bytes), "cudaMalloc(ctl)");
        CK(sb.dense_off.reserve(8 * (n + W)), "cudaMalloc(dense_off)");
        if (code_size) CK(sb.cs.reserve(n), "cudaMalloc(code_size)");
    }
    uint8_t* mp = (uint8_t*)sb.meta.p;
    auto carve = [&](size_t bytes) {
        uint8_t* r = mp;
        mp += up8(bytes);
        return r;
    };
    uint64_t* h_in_off = (uint64_t*)carve(8 * (n + 1));
    uint64_t* h_slots = (uint64_t*)carve(8 * (n + 1));
    uint32_t* h_order = (uint32_t*)carve(4 * n);
    uint16_t* h_win_of = (uint16_t*)carve(2 * n);
    uint32_t* h_win_count = (uint32_t*)carve(4 * W);
    uint32_t* h_avail = (uint32_t*)carve(4 * W);
    uint8_t* h_zero = carve(ctl_bytes);
    uint8_t* h_cs = carve(n);
    uint64_t* h_dense_off = (uint64_t*)carve(8 * (n + W));
    uint32_t* h_status = (uint32_t*)carve(4 * n);
    uint32_t* h_detail = (uint32_t*)carve(4 * n);
    volatile uint32_t* h_flags = (volatile uint32_t*)sb.flags.p;
    uint32_t* d_flags = nullptr;
    CK(cudaHostGetDevicePointer((void**)&d_flags, sb.flags.p, 0), "cudaHostGetDevicePointer");
    for (size_t w = 0; w <= W; w++) h_flags[w] = 0;
    memset(h_zero, 0, ctl_bytes);
    for (size_t w = 0; w < W; w++) h_avail[w] = (uint32_t)(w + 1);
    unsigned long long* d_queue = (unsigned long long*)sb.ctl.p;
    uint32_t* d_avail = (uint32_t*)((uint8_t*)sb.ctl.p + 64);
    uint32_t* d_done = (uint32_t*)((uint8_t*)sb.ctl.p + 128);

    HostTrace tr;
    tr.begin(W, ep.in);
    auto fail = [&](int rc) {
        // the kernel may still be waiting for input that will not come: it gives up by itself
        ep.sync_all();
        cudaGetLastError();
        return rc;
    };
#define CKS(call, what)                                               \
    do {                                                              \
        cudaError_t e_ = (call);                                      \
        if (e_ != cudaSuccess) return fail(fail_cuda(ctx, e_, what)); \
    } while (0)
    // ---- input copies, each followed by the word that tells the kernel the window is in ----
    CKS(cudaMemcpyAsync(sb.ctl.p, h_zero, ctl_bytes, cudaMemcpyHostToDevice, ep.in), "H2D ctl");
    auto copy_window = [&](size_t w) -> int {
        const uint64_t lo = in_off[wb[w]], hi = in_off[wb[w + 1]];
        tr.enqueued(w);
        if (hi > lo)
            CK(cudaMemcpyAsync((uint8_t*)sb.in.p + (lo - in_lo), in + lo, hi - lo, cudaMemcpyHostToDevice, ep.in), "H2D in");
        CK(cudaMemcpyAsync(d_avail, h_avail + w, 4, cudaMemcpyHostToDevice, ep.in), "H2D avail");
        tr.mark(w, 0, ep.in);
        return SLZW_RC_OK;
    };
    int rc;
    const size_t early = W < 2 ? W : 2;  // these copies run while the host prepares the small arrays
    for (size_t w = 0; w < early; w++)
        if ((rc = copy_window(w)) != SLZW_RC_OK) return fail(rc);
    // ---- small arrays: offsets, worst-case slots, processing order (window by window, largest
    // streams first inside a window: counting sort over 1/8-octave size classes) ----
    memcpy(h_in_off, in_off, 8 * (n + 1));
    if (code_size) memcpy(h_cs, code_size, n);
    h_slots[0] = 0;
    {
        constexpr int kClasses = 64 * 8 + 1;
        std::vector<uint32_t> count(kClasses);
        std::vector<uint16_t> cls(n);
        for (size_t w = 0; w < W; w++) {
            const uint64_t s0 = wb[w], s1 = wb[w + 1];
            h_win_count[w] = (uint32_t)(s1 - s0);
            std::fill(count.begin(), count.end(), 0u);
            for (uint64_t i = s0; i < s1; i++) {
                const uint64_t len = in_off[i + 1] - in_off[i];
                h_slots[i + 1] = h_slots[i] + ((slzw_encode_bound(params, len) + 15) & ~15ull);
                int c = 0;
                if (len) {
                    const int msb = 63 - __builtin_clzll(len);
                    c = 1 + msb * 8 + (int)(msb >= 3 ? (len >> (msb - 3)) & 7 : (len << (3 - msb)) & 7);
                }
                cls[i] = (uint16_t)c;
                count[c]++;
            }
            // descending classes: start[c] = number of streams in larger classes
            uint32_t acc = 0;
            for (int c = kClasses - 1; c >= 0; c--) {
                const uint32_t k = count[c];
                count[c] = acc;
                acc += k;
            }
            for (uint64_t i = s0; i < s1; i++) {
                const uint64_t q = s0 + count[cls[i]]++;
                h_order[q] = (uint32_t)i;
                h_win_of[q] = (uint16_t)w;
            }
        }
    }
    const uint64_t slot_bytes = h_slots[n];
    {
        std::lock_guard<std::mutex> lock(ctx->mu);
        CKS(sb.slots.reserve(slot_bytes + 16), "cudaMalloc(slots)");
        CKS(sb.dense.reserve(slot_bytes + align * n + 256), "cudaMalloc(dense)");
    }
    CKS(cudaMemcpyAsync(sb.in_off.p, h_in_off, 8 * (n + 1), cudaMemcpyHostToDevice, ep.in), "H2D in_off");
    CKS(cudaMemcpyAsync(sb.out_off.p, h_slots, 8 * (n + 1), cudaMemcpyHostToDevice, ep.in), "H2D slots");
    CKS(cudaMemcpyAsync(sb.order.p, h_order, 4 * n, cudaMemcpyHostToDevice, ep.in), "H2D order");
    CKS(cudaMemcpyAsync(sb.win_of.p, h_win_of, 2 * n, cudaMemcpyHostToDevice, ep.in), "H2D win_of");
    CKS(cudaMemcpyAsync(sb.win_count.p, h_win_count, 4 * W, cudaMemcpyHostToDevice, ep.in), "H2D win_count");
    if (code_size) CKS(cudaMemcpyAsync(sb.cs.p, h_cs, n, cudaMemcpyHostToDevice, ep.in), "H2D code_size");
    CKS(cudaEventRecord(sb.ev_meta, ep.in), "cudaEventRecord");
    // ---- the one encode launch (before the remaining copies are issued: from pageable memory
    // every one of them holds the host until its bytes are staged) ----
    {
        std::lock_guard<std::mutex> lock(ctx->mu);
        DevBatch a;
        a.in = (const uint8_t*)sb.in.p - in_lo;
        a.in_off = (const uint64_t*)sb.in_off.p;
        a.out = (uint8_t*)sb.slots.p;
        a.out_off = (const uint64_t*)sb.out_off.p;
        a.out_len = (uint64_t*)sb.out_len.p;
        a.status = (uint32_t*)sb.status.p;
        a.detail = (uint32_t*)sb.detail.p;
        a.code_size = code_size ? (const uint8_t*)sb.cs.p : nullptr;
        a.n = n;
        a.order = (const uint32_t*)sb.order.p;
        a.queue = d_queue;
        a.retry = nullptr;
        a.retry_ids = nullptr;
        a.n_dev = nullptr;
        a.dec_tables = nullptr;
        a.enc_tables = nullptr;
        a.p = *params;
        a.sc.win_of = (const uint16_t*)sb.win_of.p;
        a.sc.win_count = (const uint32_t*)sb.win_count.p;
        a.sc.done = d_done;
        a.sc.avail = d_avail;
        a.sc.host_flags = d_flags;
        a.sc.host_abort = d_flags + W;
        CKS(cudaStreamWaitEvent(ep.cmp[0], sb.ev_meta, 0), "cudaStreamWaitEvent");
        int sms = ctx->num_sms - ctx->stream_reserved_sms;
        if (sms < 1) sms = 1;
        CKS(encode_launch(a, sms, ep.cmp[0]), "encode launch");
        ctx->launches += 1;
        ctx->last_encode_ws = -1;
    }
    for (size_t w = early; w < W; w++)
        if ((rc = copy_window(w)) != SLZW_RC_OK) return fail(rc);
    // ---- windows come back: compaction as soon as the kernel raises the flag, placement in order ----
    uint64_t hbase = 0;
    bool overflow = false;
    size_t launched = 0, placed = 0;
    // worst-case position of a window in the device dense buffer
    auto dense_base = [&](size_t w) { return h_slots[wb[w]] + align * wb[w]; };
    const auto t_wait0 = std::chrono::steady_clock::now();
    auto waited_s = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_wait0).count(); };
    double last_progress = 0.0;
    while (placed < W) {
        bool progress = false;
        while (launched < W && launched < placed + kStreamRing && h_flags[launched] != 0) {
            const size_t w = launched;
            const int r = (int)(w % kStreamRing);
            const uint64_t s0 = wb[w], m = wb[w + 1] - wb[w];
            uint64_t* d_doff = (uint64_t*)sb.dense_off.p + s0 + w;
            CKS(compact_launch((const uint8_t*)sb.slots.p, (const uint64_t*)sb.out_off.p + s0,
                               (const uint64_t*)sb.out_len.p + s0, m, align, (uint8_t*)sb.dense.p + dense_base(w),
                               d_doff, ctx->num_sms, ep.cmp[1]), "compaction launch");
            ctx->launches += 2;
            CKS(cudaEventRecord(sb.ev_cmp[r], ep.cmp[1]), "cudaEventRecord");
            tr.mark(w, 1, ep.cmp[1]);
            CKS(cudaStreamWaitEvent(ep.small, sb.ev_cmp[r], 0), "cudaStreamWaitEvent");
            CKS(cudaMemcpyAsync(h_dense_off + s0 + w, d_doff, 8 * (m + 1), cudaMemcpyDeviceToHost, ep.small), "D2H dense_off");
            CKS(cudaMemcpyAsync(h_status + s0, (uint32_t*)sb.status.p + s0, 4 * m, cudaMemcpyDeviceToHost, ep.small), "D2H status");
            CKS(cudaMemcpyAsync(h_detail + s0, (uint32_t*)sb.detail.p + s0, 4 * m, cudaMemcpyDeviceToHost, ep.small), "D2H detail");
            CKS(cudaEventRecord(sb.ev_small[r], ep.small), "cudaEventRecord");
            launched++;
            progress = true;
        }
        if (placed < launched) {
            const cudaError_t q = cudaEventQuery(sb.ev_small[placed % kStreamRing]);
            if (q == cudaSuccess) {
                const size_t w = placed;
                const uint64_t s0 = wb[w], m = wb[w + 1] - wb[w];
                const uint64_t* doff = h_dense_off + s0 + w;
                const uint64_t tot = doff[m];
                for (uint64_t i = 1; i <= m; i++) out_off[s0 + i] = hbase + doff[i];
                memcpy(status + s0, h_status + s0, 4 * m);
                memcpy(detail + s0, h_detail + s0, 4 * m);
                if (deferred) {  // the bytes stay in sb.dense until run_host_dense_finish
                    ctx->deferred.push_back({dense_base(w), tot});
                } else {
                    if (hbase + tot > out_cap) overflow = true;
                    if (!overflow && tot)
                        CKS(cudaMemcpyAsync(out_dense + hbase, (uint8_t*)sb.dense.p + dense_base(w), tot,
                                            cudaMemcpyDeviceToHost, ep.dense), "D2H dense");
                }
                tr.placed(w);
                tr.mark(w, 2, ep.dense);
                hbase += tot;
                placed++;
                progress = true;
            } else if (q != cudaErrorNotReady) {
                return fail(fail_cuda(ctx, q, "cudaEventQuery"));
            }
        }
        if (progress) {
            last_progress = waited_s();
        } else {
            if (h_flags[W] != 0 || waited_s() - last_progress > 30.0) {
                snprintf(ctx->err, sizeof ctx->err, "streaming encode stalled (window %zu of %zu)", placed, W);
                return fail(SLZW_RC_CUDA);
            }
            std::this_thread::yield();
        }
    }
    CKS(cudaStreamSynchronize(ep.dense), "cudaStreamSynchronize");
    CKS(cudaStreamSynchronize(ep.cmp[0]), "cudaStreamSynchronize");
    CKS(cudaStreamSynchronize(ep.cmp[1]), "cudaStreamSynchronize");
#undef CKS
    if (h_flags[W] != 0) {
        snprintf(ctx->err, sizeof ctx->err, "streaming encode: a window's input did not arrive");
        return SLZW_RC_CUDA;
    }
    tr.report("encode (dense, streaming)", in_off, wb);
    if (needed) *needed = hbase;
    if (deferred) {
        ctx->deferred_total = hbase;
        ctx->deferred_base = (const uint8_t*)sb.dense.p;
    }
    if (overflow) {
        snprintf(ctx->err, sizeof ctx->err, "dense output needs %llu bytes, capacity is %llu",
                 (unsigned long long)hbase, (unsigned long long)out_cap);
        return SLZW_RC_NOMEM;
    }
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
    // a one-phase dense encode reuses the device buffer that a streaming _begin left its bytes in:
    // what that _begin left behind is discarded
    if (!deferred && ctx->deferred_base && ctx->deferred_base == (const uint8_t*)ctx->sb.dense.p) {
        ctx->deferred.clear();
        ctx->deferred_total = 0;
    }
    if (align == 0) align = 1;
    NvtxRange range("slzw encode batch, dense (host pipeline)");
    DeviceGuard guard(ctx->device);
    if (!guard.ok) return fail_cuda(ctx, cudaGetLastError(), "cudaSetDevice");
    // one launch over the whole call when its buffers fit comfortably (4.2 x the input) and nothing
    // has to touch the input on the device first
    if (ctx->enc_stream && ctx->pred_row_bytes == 0 && n <= 0xFFFFFFFFull &&
        in_off[n] - in_off[0] <= ctx->stream_max_bytes) {
        const int rc = run_host_encode_stream(ctx, params, in, in_off, n, code_size, align, out_dense, out_cap,
                                              out_off, status, detail, needed, deferred);
        if (rc != kStreamFallback) return rc;
    }
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

    HostTrace tr;
    tr.begin(chunks, ep.in);
    auto fail = [&](int rc) {
        ep.sync_all();
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
        tr.enqueued(k);
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
        tr.mark(k, 0, ep.in);
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
        tr.mark(k, 1, sc);
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
        tr.placed(k);
        tr.mark(k, 2, ep.dense);
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
    tr.report("encode (dense)", in_off, cb);
    if (needed) *needed = hbase;
    if (deferred) {
        ctx->deferred_total = hbase;
        ctx->deferred_base = (const uint8_t*)ctx->shard_dense.p;
    }
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
            CK(cudaMemcpyAsync(dst + at, ctx->deferred_base + c.dev_off, c.bytes,
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
    if (const char* e = getenv("SLZW_HOST_ENC_STREAM")) ctx->enc_stream = atoi(e) != 0;  // tuning knobs
    if (const char* e = getenv("SLZW_HOST_WINDOW_MB")) {
        const double v = atof(e);
        if (v >= 0.001) ctx->stream_window_bytes = (uint64_t)(v * 1048576.0);
    }
    if (const char* e = getenv("SLZW_HOST_STREAM_SMS")) {
        const int v = atoi(e);
        if (v >= 0 && v < prop.multiProcessorCount / 2) ctx->stream_reserved_sms = v;
    }
    if (const char* e = getenv("SLZW_HOST_ENC_HILL")) {  // tuning knob: "first_mb:second:grow:taper[:peak_mb]"
        double a, b, c, d, pk = 0;
        if (!strcmp(e, "taper")) ctx->enc_shape = 0;
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
        ctx->sb.release();
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
