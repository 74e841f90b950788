// nglod_b200 -- OctreeSDF.sdf forward with the decoder on tcgen05 tensor cores (3xTF32, TMEM accumulator).
// See sdf_tc.cuh / tc_common.cuh for the scheme.  Reference behaviour: sdf-net/lib/models/OctreeSDF.py:94-146.
#include "sdf_tc.cuh"
#include "internal.h"

namespace {

#ifndef NGLOD_FWD_GROUPS
#define NGLOD_FWD_GROUPS 3
#endif
constexpr int FWD_GROUPS = NGLOD_FWD_GROUPS;   // groups of 128 threads per CTA, one CTA per SM
constexpr int FWD_THREADS = FWD_GROUPS * TCG_THREADS;
constexpr int FWD_SMEM = TC_SMEM_BYTES(FWD_GROUPS);

template <bool HALF>
__global__ void __launch_bounds__(FWD_THREADS, 1)
sdf_forward_tc_kernel(const NetDev net, const float* __restrict__ x, const long long n, float* __restrict__ out) {
    extern __shared__ __align__(128) char smem_tc[];
    const uint32_t tmem_base = tc_prologue(net, smem_tc, FWD_GROUPS);
    TcGroup g = tc_make_group(smem_tc, FWD_GROUPS, tmem_base);
    const long long ggroup = (long long)blockIdx.x * FWD_GROUPS + (threadIdx.x >> 7);
    const long long ngroups = (long long)gridDim.x * FWD_GROUPS;
    for (long long tile = ggroup; tile * TCG_THREADS < n; tile += ngroups) {
        const long long i = tile * TCG_THREADS + g.wq * 32 + g.lane;
        const bool active = i < n;
        float px = 0.f, py = 0.f, pz = 0.f;
        if (active) { px = __ldg(x + 3 * i); py = __ldg(x + 3 * i + 1); pz = __ldg(x + 3 * i + 2); }
        const float d = tc_group_eval<HALF>(net, g, px, py, pz, active);
        if (active) out[i] = d;
    }
    tc_epilogue_free(tmem_base);
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
    auto kern = nd.half_pairs ? sdf_forward_tc_kernel<true> : sdf_forward_tc_kernel<false>;
    NGLOD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    long long grid = nglod_sm_count();
    const long long want = (n + FWD_THREADS - 1) / FWD_THREADS;
    if (want < grid) grid = want;
    kern<<<(int)grid, FWD_THREADS, FWD_SMEM, st>>>(nd, x, n, out);
    return (int)cudaGetLastError();
}

extern "C" int nglod_debug_tc_gemm(const float* A, const float* B, float* D, void* stream) {
    if (!A || !B || !D) return NGLOD_EINVAL;
    const int smem = 4 * TC_OPERAND_BYTES + 64;
    NGLOD_CUDA_TRY(cudaFuncSetAttribute(tc_gemm_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    tc_gemm_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D);
    return (int)cudaGetLastError();
}
