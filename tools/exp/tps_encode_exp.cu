// Experiment (not product): thread-per-stream LZW matcher with generation-tagged hash tables in
// global memory.  Emits u16 codes only (no bit packing).  Measures sustained probe throughput.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int LOG_SLOTS>
__global__ void tps_match(const uint8_t* __restrict__ in, const uint64_t* __restrict__ in_off,
                          const uint32_t* __restrict__ order, uint32_t n,
                          uint64_t* __restrict__ tables, uint16_t* __restrict__ codes_out,
                          const uint64_t* __restrict__ codes_off, uint32_t* __restrict__ ncodes_out,
                          unsigned long long* queue, uint32_t gen_base) {
    constexpr uint32_t SLOTS = 1u << LOG_SLOTS;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t* tbl = tables + (uint64_t)tid * SLOTS;
    uint32_t gen = gen_base;
    const int lane = threadIdx.x & 31;
    for (;;) {
        // one warp grabs 32 consecutive streams of the (size-sorted) order
        unsigned long long q = 0;
        if (lane == 0) q = atomicAdd(queue, 32ull);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= n) break;
        const uint32_t qi = (uint32_t)q + lane;
        if (qi < n) {
            const uint32_t sid = order[qi];
            const uint8_t* src = in + in_off[sid];
            const uint32_t len = (uint32_t)(in_off[sid + 1] - in_off[sid]);
            uint16_t* out = codes_out + codes_off[sid];
            uint32_t nc = 0;
            gen++;
            uint32_t next_code = 258;
            if (len > 0) {
                uint32_t pw = (uint32_t)src[0] << 20;
                for (uint32_t i = 1; i < len; i++) {
                    const uint32_t k = __ldg(src + i);
                    const uint32_t key = ((pw >> 12) & 0xFFF00u) | k;
                    uint32_t h = (pw * 0x9E3779B1u + k * 0x85EBCA6Bu) >> (32 - LOG_SLOTS);
                    uint64_t s;
                    bool hit = false;
                    for (;;) {
                        s = tbl[h];
                        if ((uint32_t)(s >> 32) != gen) break;              // empty (stale generation)
                        if (((uint32_t)s & 0xFFFFFu) == key) { hit = true; break; }
                        h = (h + 1) & (SLOTS - 1);
                    }
                    if (hit) {
                        pw = (uint32_t)s;
                    } else {
                        const uint32_t idx = next_code++;
                        tbl[h] = ((uint64_t)gen << 32) | (idx << 20) | key;
                        out[nc++] = (uint16_t)(pw >> 20);
                        pw = k << 20;
                        if (idx == 4094) {  // TIFF reset
                            out[nc++] = 256;
                            next_code = 258;
                            gen++;
                        }
                    }
                }
                out[nc++] = (uint16_t)(pw >> 20);
            }
            ncodes_out[sid] = nc;
        }
    }
}

extern "C" __attribute__((visibility("default")))
int tps_run(const uint8_t* in, const uint64_t* in_off, const uint32_t* order, uint32_t n,
            uint64_t* tables, uint16_t* codes_out, const uint64_t* codes_off, uint32_t* ncodes_out,
            unsigned long long* queue, uint32_t gen_base, int blocks, int threads, int log_slots,
            void* stream) {
    cudaMemsetAsync(queue, 0, 8, (cudaStream_t)stream);
    if (log_slots == 13)
        tps_match<13><<<blocks, threads, 0, (cudaStream_t)stream>>>(in, in_off, order, n, tables, codes_out, codes_off, ncodes_out, queue, gen_base);
    else if (log_slots == 12)
        tps_match<12><<<blocks, threads, 0, (cudaStream_t)stream>>>(in, in_off, order, n, tables, codes_out, codes_off, ncodes_out, queue, gen_base);
    else
        tps_match<14><<<blocks, threads, 0, (cudaStream_t)stream>>>(in, in_off, order, n, tables, codes_out, codes_off, ncodes_out, queue, gen_base);
    return (int)cudaGetLastError();
}
