// Experiment, not part of libslzw.so (compiled only with -DSLZW_EXP_LANES, tools/build_variants.sh):
// one LANE per stream instead of one warp per stream.  Result: profiles/r02_encode_notes.md.
// Included into the namespace slzw of encode_kernels.cu.
template <int TILE, int TWARPS, int SWARPS, int NSL, int GW>
struct LaneConfig {
    using L = LaneLayout<TILE, TWARPS, SWARPS, NSL, GW>;
    static constexpr int kWarps = L::kWarps;
    static constexpr int kStreams = TWARPS + SWARPS + NSL + 32 * GW;  // per SM
    static uint32_t smem() {
        const uint32_t a = L::bytes(1024u + L::kHead), b = L::bytes(L::kHead);
        return (a > b ? a : b) + L::kHead;
    }
    static cudaError_t configure() {
        cudaError_t e = cudaFuncSetAttribute(slzw_encode_lanes_kernel<TILE, TWARPS, SWARPS, NSL, GW, false>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem());
        if (e != cudaSuccess) return e;
        return cudaFuncSetAttribute(slzw_encode_lanes_kernel<TILE, TWARPS, SWARPS, NSL, GW, true>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem());
    }
    static size_t table_bytes(int num_sms) { return (size_t)num_sms * GW * 32u * (kSlots * 4u); }
    static cudaError_t launch(const DevBatch& a, int num_sms, cudaStream_t stream) {
        const uint64_t ctas = (a.n + kStreams - 1) / kStreams;
        const int grid = (int)(ctas < (uint64_t)num_sms ? ctas : (uint64_t)num_sms);
        if (a.p.flavour == SLZW_FLAVOUR_FIXED)
            slzw_encode_lanes_kernel<TILE, TWARPS, SWARPS, NSL, GW, true><<<grid, kWarps * kWarpSize, smem(), stream>>>(a, smem());
        else
            slzw_encode_lanes_kernel<TILE, TWARPS, SWARPS, NSL, GW, false><<<grid, kWarps * kWarpSize, smem(), stream>>>(a, smem());
        return cudaGetLastError();
    }
};


// {tile of the warp-per-stream warps, tensor-memory warps, shared-memory warps, lanes in shared
//  memory, warps of lanes in global memory}
using Lan10 = LaneConfig<64, 0, 0, 13, 0>;
using Lan11 = LaneConfig<64, 0, 0, 13, 1>;
using Lan12 = LaneConfig<64, 0, 0, 0, 1>;
using Lan13 = LaneConfig<64, 0, 0, 0, 2>;
using Lan14 = LaneConfig<64, 0, 0, 0, 4>;
using Lan15 = LaneConfig<64, 16, 0, 13, 0>;
using Lan16 = LaneConfig<80, 16, 12, 0, 1>;
using Lan17 = LaneConfig<80, 16, 12, 0, 2>;
using Lan18 = LaneConfig<80, 16, 12, 0, 4>;
using Lan19 = LaneConfig<80, 16, 12, 0, 0>;

#define SLZW_LANE_CONFIGS(X) X(10, Lan10) X(11, Lan11) X(12, Lan12) X(13, Lan13) X(14, Lan14) X(15, Lan15) X(16, Lan16) X(17, Lan17) X(18, Lan18) X(19, Lan19)
