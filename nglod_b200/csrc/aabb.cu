// nglod_b200 -- standalone ray/cube kernel (the drop-in for sol_nglod.aabb).
// One launch writes x, t and hit for every ray (the reference needs a clone,
// two zero-fills and the kernel).  Ray data is [n,3] AoS, so a CTA stages its
// 256 rays through shared memory with fully coalesced 4-byte accesses instead
// of 12-byte-strided per-thread loads.
#include "common.cuh"
#include "aabb.cuh"

namespace {
constexpr int AABB_THREADS = 256;

__global__ void __launch_bounds__(AABB_THREADS)
aabb_kernel(const float* __restrict__ ray_o, const float* __restrict__ ray_d, const long long n,
            float* __restrict__ x, float* __restrict__ t, uint8_t* __restrict__ hit) {
    __shared__ float so[AABB_THREADS * 3];
    __shared__ float sd[AABB_THREADS * 3];
    const long long nblk = (n + AABB_THREADS - 1) / AABB_THREADS;
    for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const long long base = blk * AABB_THREADS;
        const int cnt = (int)min((long long)AABB_THREADS, n - base);
        for (int e = threadIdx.x; e < cnt * 3; e += AABB_THREADS) {
            so[e] = __ldg(ray_o + base * 3 + e);
            sd[e] = __ldg(ray_d + base * 3 + e);
        }
        __syncthreads();
        AabbResult r;
        const int k = threadIdx.x;
        if (k < cnt) {
            // stride-3 word reads: gcd(3,32)=1 -> conflict-free
            r = ray_unit_cube(so[3 * k], so[3 * k + 1], so[3 * k + 2], sd[3 * k], sd[3 * k + 1], sd[3 * k + 2]);
        }
        __syncthreads();
        if (k < cnt) {
            so[3 * k] = r.x; so[3 * k + 1] = r.y; so[3 * k + 2] = r.z;
            t[base + k] = r.t;
            hit[base + k] = r.hit ? 1 : 0;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < cnt * 3; e += AABB_THREADS) x[base * 3 + e] = so[e];
        __syncthreads();
    }
}
}  // namespace

extern "C" int nglod_aabb(const float* ray_o, const float* ray_d, int64_t n, float* x, float* t, uint8_t* hit,
                          void* stream) {
    if (n < 0 || (n > 0 && (!ray_o || !ray_d || !x || !t || !hit))) return NGLOD_EINVAL;
    if (n == 0) return 0;
    const long long nblk = (n + AABB_THREADS - 1) / AABB_THREADS;
    const long long cap = (long long)nglod_sm_count() * 8;
    const int grid = (int)(nblk < cap ? nblk : cap);
    aabb_kernel<<<grid, AABB_THREADS, 0, (cudaStream_t)stream>>>(ray_o, ray_d, (long long)n, x, t, hit);
    return (int)cudaGetLastError();
}
