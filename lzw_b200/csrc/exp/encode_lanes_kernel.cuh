// Experiment, not part of libslzw.so (compiled only with -DSLZW_EXP_LANES, tools/build_variants.sh):
// one LANE per stream instead of one warp per stream.  Result: profiles/r02_encode_notes.md.
// Included into the anonymous namespace of encode_kernels.cu.
// ---- the kernel with lanes ------------------------------------------------------------------------
// Warps [0, TWARPS): one stream per warp, dictionary in tensor memory (encode_stream above).
// Warps [TWARPS, TWARPS + SWARPS): one stream per warp, dictionary in shared memory (bucket lookups).
// Then, if NSL > 0, one warp of NSL lanes with dictionaries in shared memory, then GW warps of 32
// lanes each with dictionaries in global memory (a.enc_tables, 16 KB per lane).  All take streams
// from the same queue.
//
// Shared memory: the SWARPS + NSL tables start at the first 16 KB boundary (so that
// `table | offset` needs no add); the per-warp blocks of the warp-per-stream warps and the input
// rings of the lanes fill the space in front of the first table and behind the last one.
template <int TILE, int TWARPS, int SWARPS, int NSL, int GW>
struct LaneLayout {
    static constexpr uint32_t kTable = kSlots * 4;
    static constexpr uint32_t kMisc = (uint32_t)((sizeof(EncMisc<TILE>) + 31) & ~size_t(31));
    static constexpr uint32_t kRingS = (64u * NSL + 63u) & ~63u;  // 4 chunks per lane
    static constexpr uint32_t kRingG = 32u * 32u;                  // 2 chunks per lane
    static constexpr uint32_t kHead = 64;                          // tensor-memory base address slot
    static constexpr int kWps = TWARPS + SWARPS;                   // warp-per-stream warps
    static constexpr int kLaneWarp = NSL > 0 ? 1 : 0;
    static constexpr int kItems = kWps + kLaneWarp + GW;
    static constexpr int kWarps = kItems;
    __host__ __device__ static uint32_t item_bytes(int i) {
        return i < kWps ? kMisc : (i < kWps + kLaneWarp ? kRingS : kRingG);
    }
    __host__ __device__ static uint32_t first_table(uint32_t base) { return (base + kTable - 1) & ~(kTable - 1); }
    // shared-window address of item i; `end` receives the end of the last item
    __host__ __device__ static uint32_t place(uint32_t base, int want, uint32_t* end) {
        uint32_t front = (base + 63u) & ~63u;
        const uint32_t front_hi = first_table(base);
        uint32_t back = front_hi + (SWARPS + NSL) * kTable;
        uint32_t at = 0;
        for (int i = 0; i < kItems; i++) {
            const uint32_t b = item_bytes(i);
            uint32_t where;
            if (front + b <= front_hi) {
                where = front;
                front += b;
            } else {
                where = back;
                back += b;
            }
            if (i == want) at = where;
        }
        if (end) *end = back;
        return at;
    }
    __host__ __device__ static uint32_t bytes(uint32_t base) {
        uint32_t end;
        place(base, 0, &end);
        return end - base;
    }
};

template <int TILE, int TWARPS, int SWARPS, int NSL, int GW, bool FIXED>
__global__ void __launch_bounds__(LaneLayout<TILE, TWARPS, SWARPS, NSL, GW>::kWarps * kWarpSize, 1)
slzw_encode_lanes_kernel(const DevBatch a, const uint32_t dyn_bytes) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    using L = LaneLayout<TILE, TWARPS, SWARPS, NSL, GW>;
    static_assert(TWARPS == 0 || TWARPS == 16, "tensor memory holds 16 dictionaries of 128 columns");
    const uint32_t warp = threadIdx.x / kWarpSize;
    const int lane = threadIdx.x % kWarpSize;
    const uint32_t raw = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t base = raw + L::kHead;
    if (L::bytes(base) + L::kHead > dyn_bytes) __trap();  // launch configuration and layout disagree

    uint32_t tmem_base = 0;
    if constexpr (TWARPS > 0) {
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(raw), "r"(512u));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_raw);
    }

    if (warp < (uint32_t)L::kWps) {
        if constexpr (L::kWps > 0) {
            EncMisc<TILE>& S = *reinterpret_cast<EncMisc<TILE>*>(smem_raw + (L::place(base, (int)warp, nullptr) - raw));
            const bool tmem_warp = warp < (uint32_t)TWARPS;
            uint32_t tb;
            uint32_t* table = nullptr;
            if (tmem_warp) {
                tb = tmem_base + (((warp & 3u) * 32u) << 16) + (warp >> 2) * 128u;
            } else {
                tb = L::first_table(base) + (warp - TWARPS) * L::kTable;
                table = reinterpret_cast<uint32_t*>(smem_raw + (tb - raw));
            }
            for (;;) {
                unsigned long long q = 0;
                if (lane == 0) q = atomicAdd(a.queue, 1ull);
                q = __shfl_sync(kFullMask, q, 0);
                if (q >= a.n) break;
                const uint32_t sid = a.order ? a.order[q] : (uint32_t)q;
                encode_stream<TILE, FIXED, (TWARPS > 0), 2>(a, sid, table, tb, tmem_warp, S, lane);
                if (lane == 0) atomicAdd(a.queue + (tmem_warp ? 4 : 5), a.in_off[sid + 1] - a.in_off[sid]);
            }
        }
    } else if (NSL > 0 && warp == (uint32_t)L::kWps) {
        const bool enabled = lane < NSL;
        const uint32_t tbl = L::first_table(base) + (SWARPS + (enabled ? (uint32_t)lane : 0u)) * L::kTable;
        const uint32_t ring = L::place(base, L::kWps, nullptr) + 64u * (uint32_t)lane;
        encode_lanes<FIXED, false, 4>(a, tbl, 0u, ring, enabled, lane);
    } else {
        if constexpr (GW > 0) {
            const uint32_t gw = warp - L::kWps - L::kLaneWarp;
            const uint64_t tbl = reinterpret_cast<uint64_t>(a.enc_tables) +
                                 (((uint64_t)blockIdx.x * GW + gw) * kWarpSize + (uint64_t)lane) * L::kTable;
            const uint32_t ring = L::place(base, L::kWps + L::kLaneWarp + (int)gw, nullptr) + 32u * (uint32_t)lane;
            encode_lanes<FIXED, true, 2>(a, (uint32_t)tbl, (uint32_t)(tbl >> 32), ring, true, lane);
        }
    }

    if constexpr (TWARPS > 0) {
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
        if (warp == 0)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512u));
    }
}

