// slzw_multi.cu -- one host batch over several GPUs of one box (include/slzw.h, slzw_multi_*).
//
// Streams are independent, so a batch shards by stream: contiguous ranges balanced by uncompressed
// bytes, one context and one worker thread per device, NO device-side exchange; only sizes and
// statuses are gathered, on the host (BASELINE north_star: "batches shard naturally across the 8
// GPUs of one box by stream, with no NCCL collective ... only host-side size gathering").  The
// reference's API is one call per job (lzw/src/encoder.rs:479-487, lib.rs:51-91); this is that
// one call for a box.
#include <cstdio>
#include <cstring>
#include <new>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/slzw.h"

struct slzw_multi {
    std::vector<slzw_ctx*> ctx;
    std::vector<int> dev;
    char err[320] = {0};
};

namespace {

// Runs fn(k) for every device k on its own thread; returns the first non-OK code and keeps its text.
template <class F>
int for_each_device(slzw_multi* m, F fn) {
    const size_t nd = m->ctx.size();
    std::vector<int> rc(nd, SLZW_RC_OK);
    std::vector<std::thread> th;
    th.reserve(nd);
    for (size_t k = 1; k < nd; k++) th.emplace_back([&, k] { rc[k] = fn(k); });
    rc[0] = fn(0);
    for (auto& t : th) t.join();
    for (size_t k = 0; k < nd; k++)
        if (rc[k] != SLZW_RC_OK) {
            snprintf(m->err, sizeof m->err, "device %d: %s", m->dev[k], slzw_last_error(m->ctx[k]));
            return rc[k];
        }
    return SLZW_RC_OK;
}

bool batch_ok(slzw_multi* m, const slzw_params* params, const slzw_batch* b) {
    if (!m || !params || !b || (b->n && (!b->in_off || !b->out_off || !b->out_len || !b->status || !b->detail))) {
        if (m) snprintf(m->err, sizeof m->err, "invalid params or batch");
        return false;
    }
    return true;
}

slzw_batch sub_batch(const slzw_batch* b, uint64_t lo, uint64_t hi) {
    // offsets stay absolute (the host entry points take them that way)
    slzw_batch s = *b;
    s.in_off = b->in_off + lo;
    s.out_off = b->out_off + lo;
    s.out_len = b->out_len + lo;
    s.status = b->status + lo;
    s.detail = b->detail + lo;
    s.code_size = b->code_size ? b->code_size + lo : nullptr;
    s.n = hi - lo;
    return s;
}

}  // namespace

extern "C" {

void slzw_partition_streams(const uint64_t* weight_off, uint64_t n, int parts, uint64_t* bounds) {
    if (!bounds || parts <= 0) return;
    bounds[0] = 0;
    const uint64_t total = (n && weight_off) ? weight_off[n] - weight_off[0] : 0;
    uint64_t i = 0;
    for (int p = 1; p < parts; p++) {
        // first stream whose start is at or beyond p / parts of the bytes (128-bit product: no overflow)
        const uint64_t target = weight_off ? weight_off[0] + (uint64_t)(((unsigned __int128)total * (unsigned)p) / (unsigned)parts) : 0;
        while (i < n && weight_off[i] < target) i++;
        bounds[p] = i;
    }
    bounds[parts] = n;
}

int slzw_multi_create(const int* devices, int n_devices, slzw_multi** out) {
    if (!out) return SLZW_RC_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return SLZW_RC_NO_DEVICE;
    slzw_multi* m = new (std::nothrow) slzw_multi;
    if (!m) return SLZW_RC_NOMEM;
    if (!devices) {
        if (n_devices <= 0 || n_devices > count) n_devices = count;
        for (int d = 0; d < n_devices; d++) m->dev.push_back(d);
    } else {
        if (n_devices <= 0) {
            delete m;
            return SLZW_RC_INVALID;
        }
        m->dev.assign(devices, devices + n_devices);
    }
    for (int d : m->dev) {
        slzw_ctx* c = nullptr;
        const int rc = slzw_create(d, &c);
        if (rc != SLZW_RC_OK) {
            for (slzw_ctx* x : m->ctx) slzw_destroy(x);
            delete m;
            return rc;
        }
        m->ctx.push_back(c);
    }
    *out = m;
    return SLZW_RC_OK;
}

void slzw_multi_destroy(slzw_multi* m) {
    if (!m) return;
    for (slzw_ctx* c : m->ctx) slzw_destroy(c);
    delete m;
}

int slzw_multi_device_count(const slzw_multi* m) { return m ? (int)m->ctx.size() : 0; }
const char* slzw_multi_last_error(const slzw_multi* m) { return m ? m->err : "null context"; }

uint64_t slzw_multi_kernel_launches(const slzw_multi* m) {
    uint64_t t = 0;
    if (m)
        for (slzw_ctx* c : m->ctx) t += slzw_kernel_launches(c);
    return t;
}

int slzw_multi_encode_batch_host(slzw_multi* m, const slzw_params* params, const slzw_batch* b) {
    if (!batch_ok(m, params, b)) return SLZW_RC_INVALID;
    if (b->n == 0) return SLZW_RC_OK;
    const int nd = (int)m->ctx.size();
    std::vector<uint64_t> bounds(nd + 1);
    slzw_partition_streams(b->in_off, b->n, nd, bounds.data());
    return for_each_device(m, [&](size_t k) -> int {
        if (bounds[k + 1] == bounds[k]) return SLZW_RC_OK;
        const slzw_batch s = sub_batch(b, bounds[k], bounds[k + 1]);
        return slzw_encode_batch_host(m->ctx[k], params, &s);
    });
}

int slzw_multi_decode_batch_host(slzw_multi* m, const slzw_params* params, const slzw_batch* b) {
    if (!batch_ok(m, params, b)) return SLZW_RC_INVALID;
    if (b->n == 0) return SLZW_RC_OK;
    const int nd = (int)m->ctx.size();
    std::vector<uint64_t> bounds(nd + 1);
    slzw_partition_streams(b->out_off, b->n, nd, bounds.data());  // balance by decoded bytes
    return for_each_device(m, [&](size_t k) -> int {
        if (bounds[k + 1] == bounds[k]) return SLZW_RC_OK;
        const slzw_batch s = sub_batch(b, bounds[k], bounds[k + 1]);
        return slzw_decode_batch_host(m->ctx[k], params, &s);
    });
}

int slzw_multi_encode_batch_host_dense(slzw_multi* m, const slzw_params* params, const uint8_t* in,
                                       const uint64_t* in_off, uint64_t n, const uint8_t* code_size,
                                       uint64_t align, uint8_t* out_dense, uint64_t out_cap,
                                       uint64_t* out_off, uint32_t* status, uint32_t* detail,
                                       uint64_t* needed) {
    if (!m || !params || !in_off || !out_off || !status || !detail) {
        if (m) snprintf(m->err, sizeof m->err, "invalid arguments");
        return SLZW_RC_INVALID;
    }
    out_off[0] = 0;
    if (needed) *needed = 0;
    if (n == 0) return SLZW_RC_OK;
    const int nd = (int)m->ctx.size();
    std::vector<uint64_t> bounds(nd + 1);
    slzw_partition_streams(in_off, n, nd, bounds.data());
    // phase 1: every device encodes and compacts its shard; sizes come back, bytes stay put
    std::vector<uint64_t> total(nd, 0);
    int rc = for_each_device(m, [&](size_t k) -> int {
        const uint64_t lo = bounds[k], cnt = bounds[k + 1] - lo;
        if (cnt == 0) return SLZW_RC_OK;
        // out_off[lo .. lo + cnt] receives shard-relative offsets; entry lo is shared with the
        // previous shard's last entry, so the shard writes into a private copy of that one
        std::vector<uint64_t> rel(cnt + 1);
        const int r = slzw_encode_batch_host_dense_begin(m->ctx[k], params, in, in_off + lo, cnt,
                                                         code_size ? code_size + lo : nullptr, align,
                                                         rel.data(), status + lo, detail + lo, &total[k]);
        if (r == SLZW_RC_OK) memcpy(out_off + lo + 1, rel.data() + 1, sizeof(uint64_t) * cnt);
        return r;
    });
    if (rc != SLZW_RC_OK) return rc;
    // host-side size gathering: where every shard starts in the dense buffer
    std::vector<uint64_t> base(nd + 1, 0);
    for (int k = 0; k < nd; k++) base[k + 1] = base[k] + total[k];
    if (needed) *needed = base[nd];
    if (base[nd] > out_cap || (base[nd] && !out_dense)) {
        snprintf(m->err, sizeof m->err, "dense output needs %llu bytes, capacity is %llu",
                 (unsigned long long)base[nd], (unsigned long long)out_cap);
        return SLZW_RC_NOMEM;
    }
    // phase 2: every device copies its shard to its place; offsets become absolute
    return for_each_device(m, [&](size_t k) -> int {
        const uint64_t lo = bounds[k], cnt = bounds[k + 1] - lo;
        if (cnt == 0) return SLZW_RC_OK;
        for (uint64_t i = 1; i <= cnt; i++) out_off[lo + i] += base[k];
        return slzw_encode_batch_host_dense_finish(m->ctx[k], out_dense + base[k], total[k]);
    });
}

}  // extern "C"
