// nglod_b200 -- measurement probe: what the L2 -> SM path delivers for the SDF kernels' access pattern.
//
// The inference kernels read, per query, the 8 corner lines (128 B each) of one random cell of a prefix-summed grid
// that is resident in L2 (35 MB at lod 4), 8 lanes x LDG.128 per line.  HBM is not the roof for that (DRAM traffic is
// the grid once per launch); the roof is whatever the L2 slices + crossbar deliver to the SMs for scattered 128-byte
// lines.  MEASURED_PEAKS.json has no such number, so bench.py measures it on the box with this kernel: the SAME
// addresses the forward kernel generates (cell from a hash of the query index, corner offsets of a channels-last
// [S][S][S][32] fp32 grid), NO arithmetic beyond an XOR that keeps the loads alive, as many loads in flight as the
// register file allows.  roofline.peak for the gather-bound kernels is the best figure of this probe.
#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t probe_hash(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// CELLS = cells (queries) in flight per 8-lane sub-warp: 8 * CELLS independent LDG.128 per lane before the first use.
template <int CELLS>
__global__ void __launch_bounds__(512)
probe_gather_kernel(const uint4* __restrict__ grid, const int R, const long long n_queries, const uint32_t seed,
                    uint32_t* __restrict__ sink) {
    const int S = R + 1;
    const int c = threadIdx.x & 7;
    const long long sub = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const long long nsub = ((long long)gridDim.x * blockDim.x) >> 3;
    uint32_t acc = 0;
    for (long long q0 = sub * CELLS; q0 < n_queries; q0 += nsub * CELLS) {
        uint4 v[CELLS][8];
#pragma unroll
        for (int k = 0; k < CELLS; ++k) {
            const uint32_t h = probe_hash((uint32_t)(q0 + k) * 2654435761u + seed);
            const int x0 = (int)((h & 0x3ffu) * (uint32_t)R >> 10);
            const int y0 = (int)(((h >> 10) & 0x3ffu) * (uint32_t)R >> 10);
            const int z0 = (int)(((h >> 20) & 0x3ffu) * (uint32_t)R >> 10);
            const uint4* g = grid + ((size_t)((z0 * S + y0) * S + x0) * 8 + c);
            const int dx = 8, dy = S * 8, dz = S * S * 8;
            v[k][0] = __ldg(g);           v[k][1] = __ldg(g + dx);
            v[k][2] = __ldg(g + dy);      v[k][3] = __ldg(g + dy + dx);
            v[k][4] = __ldg(g + dz);      v[k][5] = __ldg(g + dz + dx);
            v[k][6] = __ldg(g + dz + dy); v[k][7] = __ldg(g + dz + dy + dx);
        }
#pragma unroll
        for (int k = 0; k < CELLS; ++k)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc ^= v[k][j].x ^ v[k][j].y ^ v[k][j].z ^ v[k][j].w;
    }
    if (acc == 0x9e3779b9u) sink[0] = acc;      // practically never true: keeps the loads without a store per thread
}

// Same byte count, no cell structure: every 8-lane sub-warp reads ONE random 128-byte line per load.
template <int LINES>
__global__ void __launch_bounds__(512)
probe_lines_kernel(const uint4* __restrict__ buf, const long long n_lines_buf, const long long n_reads, const uint32_t seed,
                   uint32_t* __restrict__ sink) {
    const int c = threadIdx.x & 7;
    const long long sub = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const long long nsub = ((long long)gridDim.x * blockDim.x) >> 3;
    uint32_t acc = 0;
    for (long long r0 = sub * LINES; r0 < n_reads; r0 += nsub * LINES) {
        uint4 v[LINES];
#pragma unroll
        for (int k = 0; k < LINES; ++k) {
            const uint32_t h = probe_hash((uint32_t)(r0 + k) * 2654435761u + seed);
            const long long line = (long long)(((unsigned long long)h * (unsigned long long)n_lines_buf) >> 32);
            v[k] = __ldg(buf + line * 8 + c);
        }
#pragma unroll
        for (int k = 0; k < LINES; ++k) acc ^= v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
    }
    if (acc == 0x9e3779b9u) sink[0] = acc;
}


// The backward's address stream: per query 8 corner lines, 8 lanes x red.global.add.v4.f32 per line, no arithmetic.
// What the L2 atomic units take is the roof of the grid-gradient scatter.
__global__ void __launch_bounds__(512)
probe_scatter_kernel(float* __restrict__ grid, const int R, const long long n_queries, const uint32_t seed) {
    const int S = R + 1;
    const int c = threadIdx.x & 7;
    const long long sub = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const long long nsub = ((long long)gridDim.x * blockDim.x) >> 3;
    for (long long q = sub; q < n_queries; q += nsub) {
        const uint32_t h = probe_hash((uint32_t)q * 2654435761u + seed);
        const int x0 = (int)((h & 0x3ffu) * (uint32_t)R >> 10);
        const int y0 = (int)(((h >> 10) & 0x3ffu) * (uint32_t)R >> 10);
        const int z0 = (int)(((h >> 20) & 0x3ffu) * (uint32_t)R >> 10);
        float* g = grid + ((size_t)((z0 * S + y0) * S + x0) * 32 + 4 * c);
        const int dx = 32, dy = S * 32, dz = S * S * 32;
        const float v = __uint_as_float((h & 0x007fffffu) | 0x3f000000u) * 1e-6f;
        red_add_v4(g, v, v, v, v);                     red_add_v4(g + dx, v, v, v, v);
        red_add_v4(g + dy, v, v, v, v);                red_add_v4(g + dy + dx, v, v, v, v);
        red_add_v4(g + dz, v, v, v, v);                red_add_v4(g + dz + dx, v, v, v, v);
        red_add_v4(g + dz + dy, v, v, v, v);           red_add_v4(g + dz + dy + dx, v, v, v, v);
    }
}

// Launch shape: as many 512-thread CTAs per SM as fit next to `smem` bytes of (unused) dynamic shared memory each -- the
// carve-out decides how much of the 228 KB is left to L1, which is where the in-flight lines of a gather live.
template <typename K>
int probe_grid(K kern, int smem, int ctas_per_sm_cap) {
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 512, smem) != cudaSuccess || per_sm < 1) return -1;
    if (ctas_per_sm_cap > 0 && per_sm > ctas_per_sm_cap) per_sm = ctas_per_sm_cap;
    return nglod_sm_count() * per_sm;
}
#define PROBE_LAUNCH(KERN, ...)                                                      \
    do {                                                                             \
        const int g_ = probe_grid(KERN, smem_bytes, ctas_per_sm);                    \
        if (g_ < 1) return NGLOD_EINVAL;                                             \
        KERN<<<g_, 512, smem_bytes, st>>>(__VA_ARGS__);                              \
    } while (0)

}  // namespace

extern "C" int nglod_probe_gather(const void* buf, int32_t grid_res, int64_t n_queries, int32_t in_flight,
                                  int32_t structured, int32_t smem_bytes, int32_t ctas_per_sm, uint32_t seed,
                                  uint32_t* sink, void* stream) {
    if (!buf || !sink || grid_res < 1 || grid_res > 256 || n_queries < 0) return NGLOD_EINVAL;
    if (smem_bytes < 0 || smem_bytes > 232448 || ctas_per_sm < 0) return NGLOD_EINVAL;
    if ((reinterpret_cast<uintptr_t>(buf) & 127u) != 0) return NGLOD_EINVAL;
    if (n_queries == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const uint4* g = reinterpret_cast<const uint4*>(buf);
    const long long S = grid_res + 1;
    if (structured) {
        switch (in_flight) {
            case 1: PROBE_LAUNCH(probe_gather_kernel<1>, g, grid_res, n_queries, seed, sink); break;
            case 2: PROBE_LAUNCH(probe_gather_kernel<2>, g, grid_res, n_queries, seed, sink); break;
            case 3: PROBE_LAUNCH(probe_gather_kernel<3>, g, grid_res, n_queries, seed, sink); break;
            default: return NGLOD_EINVAL;
        }
    } else {
        const long long lines = S * S * S, reads = (long long)n_queries * 8;
        switch (in_flight) {
            case 1: PROBE_LAUNCH(probe_lines_kernel<8>, g, lines, reads, seed, sink); break;
            case 2: PROBE_LAUNCH(probe_lines_kernel<16>, g, lines, reads, seed, sink); break;
            case 3: PROBE_LAUNCH(probe_lines_kernel<24>, g, lines, reads, seed, sink); break;
            default: return NGLOD_EINVAL;
        }
    }
    return (int)cudaGetLastError();
}

extern "C" int nglod_probe_scatter(void* buf, int32_t grid_res, int64_t n_queries, int32_t smem_bytes, int32_t ctas_per_sm,
                                   uint32_t seed, void* stream) {
    if (!buf || grid_res < 1 || grid_res > 256 || n_queries < 0) return NGLOD_EINVAL;
    if (smem_bytes < 0 || smem_bytes > 232448) return NGLOD_EINVAL;
    if ((reinterpret_cast<uintptr_t>(buf) & 127u) != 0) return NGLOD_EINVAL;
    if (n_queries == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (ctas_per_sm < 0) {          // exactly -ctas_per_sm CTAs: is the limit per SM or chip-wide?
        probe_scatter_kernel<<<-ctas_per_sm, 512, 0, st>>>(reinterpret_cast<float*>(buf), grid_res, n_queries, seed);
        return (int)cudaGetLastError();
    }
    PROBE_LAUNCH(probe_scatter_kernel, reinterpret_cast<float*>(buf), grid_res, n_queries, seed);
    return (int)cudaGetLastError();
}
