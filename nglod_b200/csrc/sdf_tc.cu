// nglod_b200 -- OctreeSDF.sdf forward with the decoder on tcgen05 tensor cores (3xTF32, TMEM accumulator).
// See sdf_tc.cuh / tc_common.cuh for the scheme.  Reference behaviour: sdf-net/lib/models/OctreeSDF.py:94-146.
#include "sdf_tc.cuh"
#include "internal.h"

namespace {

#ifndef NGLOD_FWD_GROUPS
#define NGLOD_FWD_GROUPS 3          // multi-LOD gather: groups of 128 threads per CTA, one CTA per SM
#endif
#ifndef NGLOD_FWD_GROUPS_SINGLE
#define NGLOD_FWD_GROUPS_SINGLE 4   // single-grid gather: no scratch, <=128 registers -> 16 warps, all 512 TMEM columns
#endif
constexpr int fwd_groups(int mode) { return mode == TC_MULTI ? NGLOD_FWD_GROUPS : NGLOD_FWD_GROUPS_SINGLE; }
constexpr int fwd_smem(int mode) { return TC_SMEM_BYTES_W(fwd_groups(mode), tc_mode_scratch(mode)); }

template <int MODE>
__global__ void __launch_bounds__(fwd_groups(MODE) * TCG_THREADS, 1)
sdf_forward_tc_kernel(const NetDev net, const float* __restrict__ x, const long long n, float* __restrict__ out) {
    extern __shared__ __align__(128) char smem_tc[];
    constexpr int G = fwd_groups(MODE), W = tc_mode_scratch(MODE);
    if constexpr (MODE != TC_MULTI) {
        // A big batch touches every line of the grid many times, but after a cold start (new weights / flushed L2) the
        // first tiles would fetch their lines from DRAM one dependent round at a time (~10 us before the L2 is warm):
        // stream the grid into L2 at HBM speed while the prologue stages the weights.
        if (n >= (1ll << 17)) {
            const long long S = net.res[0] + 1;
            const long long bytes = MODE == TC_SINGLE_HALF ? S * S * net.res[0] * 128 : S * S * S * NGLOD_F * 4;
            const char* base = reinterpret_cast<const char*>(net.grids[0]);
            for (long long off = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 128; off < bytes;
                 off += (long long)gridDim.x * blockDim.x * 128)
                asm volatile("prefetch.global.L2 [%0];" :: "l"(base + off));
        }
    }
    const uint32_t tmem_base = tc_prologue(net, smem_tc, G, W);
    TcGroup g = tc_make_group(smem_tc, G, tmem_base, W);
    const long long ggroup = (long long)blockIdx.x * G + (threadIdx.x >> 7);
    const long long ngroups = (long long)gridDim.x * G;
    // the next tile's coordinates are fetched (from DRAM) while the current tile is evaluated
    long long tile = ggroup;
    long long i = tile * TCG_THREADS + g.wq * 32 + g.lane;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (i < n) { px = __ldg(x + 3 * i); py = __ldg(x + 3 * i + 1); pz = __ldg(x + 3 * i + 2); }
    for (; tile * TCG_THREADS < n; tile += ngroups) {
        const long long i_next = i + ngroups * TCG_THREADS;
        float nx = 0.f, ny = 0.f, nz = 0.f;
        if (i_next < n) { nx = __ldg(x + 3 * i_next); ny = __ldg(x + 3 * i_next + 1); nz = __ldg(x + 3 * i_next + 2); }
        const bool active = i < n;
        const float d = tc_group_eval<MODE>(net, g, px, py, pz, active);
        if (active) out[i] = d;
        i = i_next; px = nx; py = ny; pz = nz;
    }
    tc_epilogue_free(tmem_base);
}

template <int MODE>
int launch_fwd(const NetDev& nd, const float* x, long long n, float* out, cudaStream_t st) {
    auto kern = sdf_forward_tc_kernel<MODE>;
    constexpr int threads = fwd_groups(MODE) * TCG_THREADS, smem = fwd_smem(MODE);
    static_assert(smem <= 232448, "shared memory budget");
    NGLOD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    long long grid = nglod_sm_count();
    const long long want = (n + threads - 1) / threads;
    if (want < grid) grid = want;
    kern<<<(int)grid, threads, smem, st>>>(nd, x, n, out);
    return (int)cudaGetLastError();
}

// Debug / self-test: D[128,128] = A[128,40] * B[128,40]^T through the exact operand layout, descriptors,
// 3xTF32 issue sequence and TMEM read-back the SDF kernels use.
__global__ void __launch_bounds__(128, 1)
tc_gemm_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
    extern __shared__ __align__(128) char smem_tc[];
    char* b_hi = smem_tc; char* b_lo = smem_tc + TC_OPERAND_BYTES;
    char* a_hi = smem_tc + 2 * TC_OPERAND_BYTES; char* a_lo = a_hi + TC_OPERAND_BYTES;
    char* misc = smem_tc + 4 * TC_OPERAND_BYTES;
    for (int e = threadIdx.x; e < 128 * TC_K; e += 128) {
        const int r = e / TC_K, k = e - r * TC_K;
        const uint32_t off = tc_elem_offset(r, k);
        float v = A[e], h = tf32_hi(v);
        *reinterpret_cast<float*>(a_hi + off) = h; *reinterpret_cast<float*>(a_lo + off) = v - h;
        v = B[e]; h = tf32_hi(v);
        *reinterpret_cast<float*>(b_hi + off) = h; *reinterpret_cast<float*>(b_lo + off) = v - h;
    }
    const uint32_t mbar = smem_u32(misc);
    if (threadIdx.x == 0) { mbar_init(mbar, 1); mbar_fence_init(); }
    if (threadIdx.x < 32) tmem_alloc(smem_u32(misc + 8), 128);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(misc + 8);
    if (threadIdx.x == 0) {
        tc_issue_tile(tmem_base, smem_u32(a_hi), smem_u32(a_lo), smem_u32(b_hi), smem_u32(b_lo));
        tc_commit(mbar);
    }
    mbar_wait(mbar, 0);
    tc_fence_after_sync();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int cb = 0; cb < 4; ++cb) {
        float v[32];
        tmem_ld32(trow + cb * 32, v);
        for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * 128 + cb * 32 + j] = v[j];
    }
    tc_fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem_base, 128);
}

}  // namespace

int nglod_launch_sdf_forward_tc(const NetDev& nd, const float* x, long long n, float* out, cudaStream_t st) {
    if (nd.half_pairs) return launch_fwd<TC_SINGLE_HALF>(nd, x, n, out, st);
    if (nd.num_lods == 1) return launch_fwd<TC_SINGLE_F32>(nd, x, n, out, st);
    return launch_fwd<TC_MULTI>(nd, x, n, out, st);
}

extern "C" int nglod_debug_tc_gemm(const float* A, const float* B, float* D, void* stream) {
    if (!A || !B || !D) return NGLOD_EINVAL;
    const int smem = 4 * TC_OPERAND_BYTES + 64;
    NGLOD_CUDA_TRY(cudaFuncSetAttribute(tc_gemm_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    tc_gemm_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D);
    return (int)cudaGetLastError();
}
