// nglod_b200 -- OctreeSDF.sdf forward, finite-difference gradient (FP32 path).
// Reference behaviour: sdf-net/lib/models/OctreeSDF.py:94-155, sdf-net/lib/diffutils.py:61-70.
#include "sdf_core.cuh"
#include <cuda_fp16.h>
#include "internal.h"

namespace {

__device__ __forceinline__ void sdf_kernel_prologue(const NetDev& net, float* smem, float*& tile, int*& idx) {
    sdf_stage_weights(net, smem);
    const int warp = threadIdx.x >> 5;
    float* wbase = smem + SDF_SMEM_WARP_OFF + warp * SDF_SMEM_PER_WARP;
    for (int e = threadIdx.x & 31; e < SDF_SMEM_PER_WARP; e += 32) wbase[e] = 0.f;
    tile = wbase;
    idx = reinterpret_cast<int*>(wbase + SDF_TILE_FLOATS);
    __syncthreads();
}

// Persistent: each warp strides over batches of 32 consecutive queries.
__global__ void __launch_bounds__(SDF_THREADS, 2)
sdf_forward_kernel(const NetDev net, const float* __restrict__ x, const long long n, float* __restrict__ out) {
    extern __shared__ __align__(16) float smem[];
    float* tile; int* idx;
    sdf_kernel_prologue(net, smem, tile, idx);
    const int lane = threadIdx.x & 31;
    const long long gwarp = (long long)blockIdx.x * SDF_WARPS + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * SDF_WARPS;
    for (long long base = gwarp * 32; base < n; base += nwarps * 32) {
        const long long i = base + lane;
        const bool active = i < n;
        float px = 0.f, py = 0.f, pz = 0.f;
        if (active) { px = __ldg(x + 3 * i); py = __ldg(x + 3 * i + 1); pz = __ldg(x + 3 * i + 2); }
        const float d = warp_sdf_eval(net, smem, tile, idx, px, py, pz, active, lane);
        if (active) out[i] = d;
    }
}

// out[i,k] = (sdf(x_i + h e_k) - sdf(x_i - h e_k)) / (2h)
__global__ void __launch_bounds__(SDF_THREADS, 2)
sdf_finitediff_kernel(const NetDev net, const float* __restrict__ x, const long long n, const float h,
                      float* __restrict__ out) {
    extern __shared__ __align__(16) float smem[];
    float* tile; int* idx;
    sdf_kernel_prologue(net, smem, tile, idx);
    const int lane = threadIdx.x & 31;
    const long long gwarp = (long long)blockIdx.x * SDF_WARPS + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * SDF_WARPS;
    const float two_h = 2.f * h;
    for (long long base = gwarp * 32; base < n; base += nwarps * 32) {
        const long long i = base + lane;
        const bool active = i < n;
        float p[3] = {0.f, 0.f, 0.f};
        if (active) { p[0] = __ldg(x + 3 * i); p[1] = __ldg(x + 3 * i + 1); p[2] = __ldg(x + 3 * i + 2); }
        float g[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float q[3] = {p[0], p[1], p[2]};
            q[k] = p[k] + h;
            const float dp = warp_sdf_eval(net, smem, tile, idx, q[0], q[1], q[2], active, lane);
            q[k] = p[k] - h;
            const float dm = warp_sdf_eval(net, smem, tile, idx, q[0], q[1], q[2], active, lane);
            g[k] = (dp - dm) / two_h;
        }
        if (active) { out[3 * i] = g[0]; out[3 * i + 1] = g[1]; out[3 * i + 2] = g[2]; }
    }
}

// out[i, 0..31] = summed interpolated features (no decoder); no weights are staged.
__global__ void __launch_bounds__(SDF_THREADS, 2)
sdf_features_kernel(const NetDev net, const float* __restrict__ x, const long long n, float* __restrict__ out) {
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* tile = smem + SDF_SMEM_WARP_OFF + warp * SDF_SMEM_PER_WARP;
    int* idx = reinterpret_cast<int*>(tile + SDF_TILE_FLOATS);
    for (int e = lane; e < SDF_SMEM_PER_WARP; e += 32) tile[e] = 0.f;
    __syncwarp();
    const long long gwarp = (long long)blockIdx.x * SDF_WARPS + warp;
    const long long nwarps = (long long)gridDim.x * SDF_WARPS;
    for (long long base = gwarp * 32; base < n; base += nwarps * 32) {
        const long long i = base + lane;
        const bool active = i < n;
        float px = 0.f, py = 0.f, pz = 0.f;
        if (active) { px = __ldg(x + 3 * i); py = __ldg(x + 3 * i + 1); pz = __ldg(x + 3 * i + 2); }
        warp_gather_tile(net, px, py, pz, active, tile, idx, lane);
        if (active) {
#pragma unroll
            for (int k4 = 0; k4 < NGLOD_F / 4; ++k4)
                *reinterpret_cast<float4*>(out + i * NGLOD_F + 4 * k4) =
                    *reinterpret_cast<const float4*>(tile + lane * NGLOD_KPAD + 4 * k4);
        }
        __syncwarp();
    }
}

// fp32 channels-last grid -> fp16 x-pair lines (see nglod_pack_grid_fp16 in the header).  One thread per 16-byte chunk:
// chunk c of line (z, y, x0) = {corner x0 ch 4c..4c+3, corner x0+1 ch 4c..4c+3}, round-to-nearest-even.
__global__ void __launch_bounds__(256)
pack_grid_fp16_kernel(const float* __restrict__ grid, const int R, const long long n_chunks, uint4* __restrict__ dst) {
    const int S = R + 1;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_chunks; e += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(e & 7);
        const long long line = e >> 3;
        const int x0 = (int)(line % R);
        const long long zy = line / R;                         // z * S + y
        const float* p0 = grid + ((zy * S + x0) * NGLOD_F + 4 * c);
        const float4 a = ldg_f4(p0), b = ldg_f4(p0 + NGLOD_F);
        const __half2 a01 = __floats2half2_rn(a.x, a.y), a23 = __floats2half2_rn(a.z, a.w);
        const __half2 b01 = __floats2half2_rn(b.x, b.y), b23 = __floats2half2_rn(b.z, b.w);
        uint4 o;
        o.x = *reinterpret_cast<const uint32_t*>(&a01); o.y = *reinterpret_cast<const uint32_t*>(&a23);
        o.z = *reinterpret_cast<const uint32_t*>(&b01); o.w = *reinterpret_cast<const uint32_t*>(&b23);
        dst[e] = o;
    }
}

// summed[node, 4c..4c+3] = sum_l trilinear(grid_l, node): one thread per (node, 16-byte chunk); the node's coordinates in
// a coarser grid are the exact rationals ix/k, so the weights (ix mod k)/k are exact for power-of-two ratios.
struct SummedArgs {
    int n_src;
    int R;
    int res[NGLOD_MAX_LODS];
    const float* grids[NGLOD_MAX_LODS];
};
__device__ __forceinline__ void node_axis(int i, int k, int Rl, int& i0, float& w1, bool& has1) {
    i0 = i / k;
    w1 = (float)(i - i0 * k) / (float)k;
    has1 = i0 < Rl;
}
// One block per (z, y) row of nodes (the row's y / z set-up is block-uniform, no per-thread 64-bit division: the first
// version spent 66 us on the 65^3 level, 1.1 TB/s), threads over the row's 8 S chunks.
__global__ void __launch_bounds__(256)
summed_grid_kernel(const SummedArgs a, float4* __restrict__ dst) {
    const int S = a.R + 1;
    const int iz = blockIdx.x / S, iy = blockIdx.x - iz * S;
    for (int t = threadIdx.x; t < S * 8; t += blockDim.x) {
        const int c = t & 7, ix = t >> 3;
        const long long e = (long long)blockIdx.x * (S * 8) + t;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int l = 0; l < a.n_src; ++l) {
            const int Rl = a.res[l], Sl = Rl + 1, k = a.R / Rl;
            int x0, y0, z0; float wx1, wy1, wz1; bool hx, hy, hz;
            node_axis(ix, k, Rl, x0, wx1, hx);
            node_axis(iy, k, Rl, y0, wy1, hy);
            node_axis(iz, k, Rl, z0, wz1, hz);
            const float wx0 = 1.f - wx1, wy0 = 1.f - wy1, wz0 = 1.f - wz1;
            const float* g = a.grids[l] + ((long long)((z0 * Sl + y0) * Sl + x0) * NGLOD_F + 4 * c);
            const int dx = hx ? NGLOD_F : 0, dy = hy ? Sl * NGLOD_F : 0, dz = hz ? Sl * Sl * NGLOD_F : 0;
            const float w00 = wx0 * wy0, w10 = wx1 * wy0, w01 = wx0 * wy1, w11 = wx1 * wy1;
            const float w[8] = {w00 * wz0, w10 * wz0, w01 * wz0, w11 * wz0, w00 * wz1, w10 * wz1, w01 * wz1, w11 * wz1};
            const int off[8] = {0, dx, dy, dy + dx, dz, dz + dx, dz + dy, dz + dy + dx};
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (w[q] != 0.f) {          // a zero-weight corner contributes nothing (and may be the clamped duplicate)
                    const float4 v = ldg_f4(g + off[q]);
                    s.x = fmaf(v.x, w[q], s.x); s.y = fmaf(v.y, w[q], s.y); s.z = fmaf(v.z, w[q], s.z); s.w = fmaf(v.w, w[q], s.w);
                }
            }
            acc.x += s.x; acc.y += s.y; acc.z += s.z; acc.w += s.w;
        }
        dst[e] = acc;
    }
}

// The octree case of the level-by-level build: dst = prolongation of `coarse` (resolution R / 2) + `own` (resolution R).
// Node weights are 0, 1/2 or 1 and the indices are shifts: ~70 instructions per 16-byte chunk where the general kernel above
// spends ~780 (runtime integer and float divisions per axis per source; ncu: 71 % issue-bound, 72 us for the 65^3 level).
__global__ void __launch_bounds__(256)
summed_grid_octree_kernel(const float4* __restrict__ coarse, const float4* __restrict__ own, const int R, float4* __restrict__ dst) {
    const int S = R + 1, Rc = R >> 1, Sc = Rc + 1;
    const int iz = blockIdx.x / S, iy = blockIdx.x - iz * S;
    const int z0 = iz >> 1, y0 = iy >> 1;
    const bool oz = iz & 1, oy = iy & 1;                 // odd: halfway between two coarse nodes (the upper one exists: iz <= R)
    const float4* row00 = coarse + (long long)(z0 * Sc + y0) * Sc * 8;
    const int dy = oy ? Sc * 8 : 0, dz = oz ? Sc * Sc * 8 : 0;
    const float wyz = (oy ? 0.5f : 1.f) * (oz ? 0.5f : 1.f);
    for (int t = threadIdx.x; t < S * 8; t += blockDim.x) {
        const int c = t & 7, ix = t >> 3;
        const long long e = (long long)blockIdx.x * (S * 8) + t;
        const float4 o = __ldg(own + e);
        const int x0 = ix >> 1;
        const bool ox = ix & 1;
        const float w = ox ? 0.5f * wyz : wyz;
        const float4* p = row00 + x0 * 8 + c;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        // same corner order as the general kernel (x fastest), zero-weight corners skipped
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const bool use = (!(q & 1) || ox) && (!(q & 2) || oy) && (!(q & 4) || oz);
            if (use) {
                const float4 v = __ldg(p + ((q & 1) ? 8 : 0) + ((q & 2) ? dy : 0) + ((q & 4) ? dz : 0));
                s.x = fmaf(v.x, w, s.x); s.y = fmaf(v.y, w, s.y); s.z = fmaf(v.z, w, s.z); s.w = fmaf(v.w, w, s.w);
            }
        }
        // the general kernel: acc = 0 + s_coarse, then acc += 1 * own
        float4 acc;
        acc.x = s.x + fmaf(o.x, 1.f, 0.f); acc.y = s.y + fmaf(o.y, 1.f, 0.f); acc.z = s.z + fmaf(o.z, 1.f, 0.f); acc.w = s.w + fmaf(o.w, 1.f, 0.f);
        dst[e] = acc;
    }
}

int launch_grid(const void* kernel, long long n) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, SDF_THREADS, SDF_SMEM_BYTES) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    const long long want = (n + SDF_THREADS - 1) / SDF_THREADS;
    long long grid = (long long)nglod_sm_count() * per_sm;
    if (want < grid) grid = want;
    return (int)(grid < 1 ? 1 : grid);
}

}  // namespace

extern "C" int nglod_sdf_forward(const nglod_net_t* net, int32_t lod, const float* x, int64_t n, float* out,
                                 void* stream) {
    if (int e = nglod_check_net(net, lod)) return e;
    if (n < 0 || (n > 0 && (!x || !out))) return NGLOD_EINVAL;
    if (n == 0) return 0;
    const NetDev nd = nglod_make_netdev_infer(net, lod);
    if (net->math_mode == NGLOD_MATH_TC3XTF32) return nglod_launch_sdf_forward_tc(nd, x, (long long)n, out, (cudaStream_t)stream);
    NGLOD_CUDA_TRY(cudaFuncSetAttribute(sdf_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SDF_SMEM_BYTES));
    const int grid = launch_grid((const void*)sdf_forward_kernel, n);
    sdf_forward_kernel<<<grid, SDF_THREADS, SDF_SMEM_BYTES, (cudaStream_t)stream>>>(nd, x, (long long)n, out);
    return (int)cudaGetLastError();
}

extern "C" int nglod_build_summed_grid(const nglod_net_t* net, int32_t lod, float* dst, void* stream) {
    if (!net || !dst || (reinterpret_cast<uintptr_t>(dst) & 15u)) return NGLOD_EINVAL;
    if (net->num_lods < 1 || net->num_lods > NGLOD_MAX_LODS || lod < 0 || lod >= net->num_lods) return NGLOD_EINVAL;
    if (net->feature_dim != NGLOD_F) return NGLOD_EUNSUPPORTED;
    SummedArgs a;
    a.n_src = lod + 1;
    a.R = net->grid_res[lod];
    for (int i = 0; i < NGLOD_MAX_LODS; ++i) { a.res[i] = 1; a.grids[i] = nullptr; }
    for (int i = 0; i <= lod; ++i) {
        if (!net->grids[i] || net->grid_res[i] < 1 || net->grid_res[i] > 256) return NGLOD_EINVAL;
        if (reinterpret_cast<uintptr_t>(net->grids[i]) & 15u) return NGLOD_EINVAL;
        if (a.R % net->grid_res[i] != 0) return NGLOD_EUNSUPPORTED;      // the grids must nest
        a.res[i] = net->grid_res[i];
        a.grids[i] = net->grids[i];
    }
    // Level by level: when the summed grid of the previous LOD is already there (net->summed[lod-1], != dst), this
    // level is its prolongation plus the LOD's own grid -- two sources instead of lod+1 (all five levels of the
    // headline model: 5 x 35 us -> ~15 us, which the trainer pays every step).
    if (lod > 0 && net->summed[lod - 1] && net->summed[lod - 1] != dst &&
        !(reinterpret_cast<uintptr_t>(net->summed[lod - 1]) & 15u)) {
        a.n_src = 2;
        a.res[0] = net->grid_res[lod - 1]; a.grids[0] = net->summed[lod - 1];
        a.res[1] = net->grid_res[lod];     a.grids[1] = net->grids[lod];
    }
    const int S = a.R + 1;
    const int threads = S * 8 < 256 ? ((S * 8 + 31) / 32) * 32 : 256;
    if (a.n_src == 2 && a.res[1] == a.R && a.res[0] * 2 == a.R && a.grids[0] != a.grids[1])
        summed_grid_octree_kernel<<<S * S, threads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(a.grids[0]),
                                                                               reinterpret_cast<const float4*>(a.grids[1]), a.R,
                                                                               reinterpret_cast<float4*>(dst));
    else
        summed_grid_kernel<<<S * S, threads, 0, (cudaStream_t)stream>>>(a, reinterpret_cast<float4*>(dst));
    return (int)cudaGetLastError();
}

extern "C" int nglod_pack_grid_fp16(const float* grid, int32_t grid_res, void* dst, void* stream) {
    if (!grid || !dst || grid_res < 1 || grid_res > 1024) return NGLOD_EINVAL;
    if ((reinterpret_cast<uintptr_t>(grid) & 15u) || (reinterpret_cast<uintptr_t>(dst) & 127u)) return NGLOD_EINVAL;
    const long long S = grid_res + 1;
    const long long n_chunks = S * S * grid_res * 8;
    long long blocks = (n_chunks + 255) / 256;
    const long long cap = (long long)nglod_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    pack_grid_fp16_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(grid, grid_res, n_chunks, static_cast<uint4*>(dst));
    return (int)cudaGetLastError();
}

extern "C" int nglod_sdf_forward_all(const nglod_net_t* net, const float* x, int64_t n, float* out, void* stream) {
    if (!net) return NGLOD_EINVAL;
    // One pass per head over the shared grids; heads differ in the LOD prefix they sum.
    for (int l = 0; l < net->num_lods; ++l) {
        if (int e = nglod_sdf_forward(net, l, x, n, out + (int64_t)l * n, stream)) return e;
    }
    return 0;
}

extern "C" int nglod_sdf_finitediff(const nglod_net_t* net, int32_t lod, const float* x, int64_t n, float h,
                                    float* out, void* stream) {
    if (int e = nglod_check_net(net, lod)) return e;
    if (n < 0 || (n > 0 && (!x || !out))) return NGLOD_EINVAL;
    if (n == 0) return 0;
    NGLOD_CUDA_TRY(cudaFuncSetAttribute(sdf_finitediff_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SDF_SMEM_BYTES));
    const NetDev nd = nglod_make_netdev_infer(net, lod, /*allow_half=*/false);      // CUDA-core kernel: fp32 lines only
    const int grid = launch_grid((const void*)sdf_finitediff_kernel, n);
    sdf_finitediff_kernel<<<grid, SDF_THREADS, SDF_SMEM_BYTES, (cudaStream_t)stream>>>(nd, x, (long long)n, h, out);
    return (int)cudaGetLastError();
}

extern "C" int nglod_sdf_features(const nglod_net_t* net, int32_t lod, const float* x, int64_t n, float* out,
                                  void* stream) {
    if (!net || net->num_lods < 1 || net->num_lods > NGLOD_MAX_LODS || lod < 0 || lod >= net->num_lods) return NGLOD_EINVAL;
    if (net->feature_dim != NGLOD_F) return NGLOD_EUNSUPPORTED;
    if (n < 0 || (n > 0 && (!x || !out))) return NGLOD_EINVAL;
    if ((reinterpret_cast<uintptr_t>(out) & 15u) != 0) return NGLOD_EINVAL;
    NetDev nd;
    nd.num_lods = lod + 1; nd.pos_invariant = 0; nd.half_pairs = 0;
    nd.w0 = nd.b0 = nd.w1 = nd.b1 = nullptr;
    for (int i = 0; i < NGLOD_MAX_LODS; ++i) {
        nd.res[i] = i <= lod ? net->grid_res[i] : 1;
        nd.grids[i] = i <= lod ? net->grids[i] : nullptr;
        // res <= 256: lod_setup works with 32-bit element offsets of (R+1)^3 * 32 (same bound as nglod_check_net)
        if (i <= lod && (!nd.grids[i] || nd.res[i] < 1 || nd.res[i] > 256 || (reinterpret_cast<uintptr_t>(nd.grids[i]) & 15u)))
            return NGLOD_EINVAL;
    }
    if (n == 0) return 0;
    NGLOD_CUDA_TRY(cudaFuncSetAttribute(sdf_features_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SDF_SMEM_BYTES));
    const int grid = launch_grid((const void*)sdf_features_kernel, n);
    sdf_features_kernel<<<grid, SDF_THREADS, SDF_SMEM_BYTES, (cudaStream_t)stream>>>(nd, x, (long long)n, out);
    return (int)cudaGetLastError();
}
