// slzw_device.cuh -- definitions shared by the sm_100a kernels and the C-ABI host layer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/slzw.h"

namespace slzw {

constexpr int kWarpSize = 32;
constexpr uint32_t kFullMask = 0xffffffffu;

// Device view of slzw_batch plus the schedule produced by the stream scheduler.
// Streaming encode (host pipeline, slzw_api.cu run_host_encode_stream): ONE encode launch covers
// the whole call while its input still arrives, window by window, over the copy engine.  Queue
// position q belongs to window win_of[q]; a warp waits until `avail` (a device word the copy
// stream writes behind each window's bytes) says that window is in, and the warp that finishes the
// last stream of a window raises the window's flag in mapped host memory, on which the host starts
// the window's compaction and copy back.  win_of == nullptr: everything is there (plain launch).
struct StreamCtl {
    const uint16_t* win_of;      // n entries
    const uint32_t* win_count;   // streams per window
    uint32_t* done;              // streams finished per window, zeroed before the launch
    const uint32_t* avail;       // number of windows whose input has arrived
    volatile uint32_t* host_flags;  // mapped pinned host memory, one word per window
    volatile uint32_t* host_abort;  // mapped pinned host memory: a wait for input timed out
};

struct DevBatch {
    const uint8_t* in;
    const uint64_t* in_off;
    uint8_t* out;             // nullptr => size-only pass (nothing is written)
    const uint64_t* out_off;  // may be nullptr when out == nullptr
    uint64_t* out_len;
    uint32_t* status;
    uint32_t* detail;
    const uint8_t* code_size;  // per-stream override or nullptr
    uint64_t n;
    const uint32_t* order;          // stream ids in processing order (largest first) or nullptr
    unsigned long long* queue;      // work-queue head, zeroed before the launch
    uint32_t* retry;                // fast decode: number of streams deferred to the exact kernel
    uint32_t* retry_ids;            // fast decode: their ids
    const uint32_t* n_dev;          // exact decode of deferred streams: stream count on the device
    uint32_t* dec_tables;           // fast decode: 16 KB table blocks of the warps without a shared-memory table
    uint8_t* enc_tables;            // encode: 16 KB dictionaries of the lanes in global memory (16 KB-aligned)
    slzw_params p;
    StreamCtl sc;                   // encode only
};

// ---- stream staging -------------------------------------------------------------------------
// Copies src[0..len) into a shared-memory tile so that byte j lands at tile[skew + j], where
// skew = (address of src) & 15.  Interior 16-byte chunks use aligned vector loads; the ragged
// head and tail use byte loads, so nothing outside [src, src+len) is ever touched.
// Returns skew.  `tile` must be 16-byte aligned and hold len + 16 bytes.
__device__ __forceinline__ uint32_t stage_tile(const uint8_t* __restrict__ src, uint32_t len,
                                               uint8_t* tile, int lane) {
    const uint32_t skew = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
    const uint8_t* base = src - skew;  // 16-byte aligned
    const uint32_t span = skew + len;  // bytes of the aligned window that matter
    const uint32_t nchunks = (span + 15u) >> 4;
    for (uint32_t c = lane; c < nchunks; c += kWarpSize) {
        const uint32_t lo = c << 4;
        if (lo >= skew && lo + 16u <= span) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + lo));
            *reinterpret_cast<uint4*>(tile + lo) = v;
        } else {
            for (uint32_t b = 0; b < 16u; b++) {
                const uint32_t o = lo + b;
                if (o >= skew && o < span) tile[o] = __ldg(base + o);
            }
        }
    }
    return skew;
}

}  // namespace slzw
