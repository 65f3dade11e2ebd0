// Dependent-chain latency of the warp-level operations the encoder's lookup step is built from
// (one warp, clock64 around 256 dependent repetitions).  nvcc -arch=sm_100a -o latency_probe latency_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define REP 256
template <int OP>
__global__ void probe(uint32_t* out, uint32_t seed) {
    __shared__ __align__(128) uint32_t sm[1024];
    const int lane = threadIdx.x;
    for (int i = lane; i < 1024; i += 32) sm[i] = (i * 7u + seed) & 1023u;
    __syncwarp();
    uint32_t x = seed + lane;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm);
    const long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < REP; r++) {
        if (OP == 0) x = x * 3u + 1u;                                           // IMAD
        if (OP == 1) x = __reduce_min_sync(0xffffffffu, x) + lane;               // CREDUX + add
        if (OP == 2) x = __ballot_sync(0xffffffffu, x & 1u) + lane;              // ISETP + VOTE + add
        if (OP == 3) x = __shfl_sync(0xffffffffu, x, x & 31) + 1u;               // SHFL
        if (OP == 4) {                                                           // ballot + flo + shfl (round-1 hit test)
            const uint32_t b = __ballot_sync(0xffffffffu, (x & 31u) == 7u) | 1u;
            x = __shfl_sync(0xffffffffu, x, 31 - __clz(b)) + lane + 1u;
        }
        if (OP == 5) {                                                           // ldmatrix bucket load
            uint32_t v;
            asm volatile("ldmatrix.sync.aligned.m8n8.x1.shared.b16 {%0}, [%1];" : "=r"(v) : "r"(sbase + ((x & 7u) << 4) + ((x & 0x38u) << 4)));
            x = v + lane;
        }
        if (OP == 6) {                                                           // scalar LDS
            uint32_t v;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(sbase + ((x & 1023u) << 2)));
            x = v;
        }
        if (OP == 7) {                                                           // LDS.128 pair + min tree (latency variant)
            uint4 a, b;
            const uint32_t ad = sbase + ((x & 127u) << 5);
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(ad));
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "r"(ad + 16));
            x = min(min(min(a.x ^ x, a.y ^ x), min(a.z ^ x, a.w ^ x)), min(min(b.x ^ x, b.y ^ x), min(b.z ^ x, b.w ^ x)));
        }
        if (OP == 8) {                                                           // uniform branch on the value
            if (__reduce_min_sync(0xffffffffu, x) > 5u) x = x * 5u + 3u; else x = x + 7u;
        }
    }
    const long long t1 = clock64();
    if (lane == 0) { out[0] = (uint32_t)((t1 - t0) / REP); out[1] = x; }
}

template <int OP> void run(const char* name, uint32_t* d) {
    uint32_t h[2];
    probe<OP><<<1, 32>>>(d, 12345u);
    probe<OP><<<1, 32>>>(d, 12345u);
    cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-52s %4u cycles per dependent repetition\n", name, h[0]);
}

// tensor memory: ld + wait as a dependent chain (address from the loaded value)
__global__ void probe_tmem(uint32_t* out) {
    __shared__ uint32_t slot;
    const int lane = threadIdx.x;
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(128u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = *(volatile uint32_t*)&slot;
    for (uint32_t c = 0; c < 128; c++) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(base + c), "r"((c * 5u + 3u) & 127u));
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    uint32_t x = lane & 127;
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < REP; r++) {
        uint32_t v;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(base + (__shfl_sync(0xffffffffu, x, 0) & 127u)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(v)::"memory");
        x = v;
    }
    long long t1 = clock64();
    if (lane == 0) out[0] = (uint32_t)((t1 - t0) / REP);
    t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < REP; r++) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(base + (r & 127)), "r"(x));
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    t1 = clock64();
    if (lane == 0) out[1] = (uint32_t)((t1 - t0) / REP);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(128u));
}

int main() {
    uint32_t* d;
    cudaMalloc(&d, 64);
    run<0>("IMAD (baseline)", d);
    run<1>("redux.sync.min + add", d);
    run<2>("setp + vote.ballot + add", d);
    run<3>("shfl.idx + add", d);
    run<4>("ballot + clz + shfl + add (round-1 hit test)", d);
    run<5>("ldmatrix.x1 + add", d);
    run<6>("ld.shared.u32", d);
    run<7>("2 x ld.shared.v4 + xor + min tree (latency variant)", d);
    run<8>("redux.sync.min + compare + branch + imad", d);
    uint32_t h[2];
    probe_tmem<<<1, 32>>>(d);
    probe_tmem<<<1, 32>>>(d);
    cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-52s %4u cycles per dependent repetition (incl. one shfl)\n", "tcgen05.ld.32x32b.x1 + wait::ld", h[0]);
    printf("%-52s %4u cycles per repetition\n", "tcgen05.st.32x32b.x1 + wait::st", h[1]);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
