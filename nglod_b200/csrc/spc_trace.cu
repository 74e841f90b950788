// nglod_b200 -- sparse OctreeSDF: in-voxel evaluation and the re-locating sphere tracer.
//
// Behavioural spec: sol-renderer/SDF.cu:218-472 (SDF::sphereTrace / getNormal) with its kernels
// sparse_grid_sample.cuh:31-109, step.cuh:31-86, ray_aabb.cuh:104-192; Python twin sdf-net/app/spc/SPCTracer.py:44-113.
// The reference runs, per march step: nonzero (sync) -> alloc -> gather kernel (read-modify-write of global
// feats_out, 8x32 scalar loads per level) -> 2 cuBLAS calls -> step kernel -> ray_aabb kernel.  Here the whole frame
// is ONE persistent kernel: per-ray state machines, the sparse gather (8 lanes per 128-byte corner row, parent chain
// walked in registers) feeding the same decoders as the dense path (tcgen05 3xTF32 tile or FP32 CUDA cores), the
// voxel re-location done inline by the ray's own lane.
#include "sdf_core.cuh"
#include "sdf_tc.cuh"
#include "sparse_core.cuh"

// re-fill attempts per round of the in-voxel tracer (see NGLOD_TRACE_REFILL_ATTEMPTS in tracer.cu).  1080p frames over the
// level-7 torus octree, lods 1-5 (profiles/exp_attempts_spc.sh): 1 attempt 1.90 / 2.28 / 2.28 / 2.76 / 3.64 ms, 4 attempts
// 2.00 / 2.40 / 2.39 / 2.84 / 3.69 ms
#ifndef NGLOD_SPC_REFILL_ATTEMPTS
#define NGLOD_SPC_REFILL_ATTEMPTS 1
#endif
// empty lanes a warp waits for before it goes back to the queue (NGLOD_TRACE_REFILL_MIN in tracer.cu).  Same frames
// (profiles/final_check.sh): 1: 1.90 / 2.28 / 2.28 / 2.76 / 3.64 ms, 8: 1.84 / 2.21 / 2.25 / 2.73 / 3.63 ms
#ifndef NGLOD_SPC_REFILL_MIN
#define NGLOD_SPC_REFILL_MIN 8
#endif

namespace {

// FP32 path: gather into the warp's [32][36] tile (see sdf_core.cuh::warp_gather_tile), then lane_decoder.
__device__ __forceinline__ float warp_sparse_eval(const SparseDev& sn, const float* sW, float* tile, int* idx, float px,
                                                  float py, float pz, int vrow, bool active, int lane) {
    const unsigned live = __ballot_sync(0xffffffffu, active);
    const int n_live = __popc(live);
    if (active) {
        idx[__popc(live & ((1u << lane) - 1u))] = lane;
        *reinterpret_cast<float4*>(tile + lane * NGLOD_KPAD + NGLOD_F) = make_float4(px, py, pz, 1.f);
    }
    __syncwarp();
    const int sub = lane >> 3, c = lane & 7;
    for (int r = 0; r * 4 < n_live; ++r) {
        const int slot = r * 4 + sub;
        const bool valid = slot < n_live;
        const int q = idx[valid ? slot : 0];
        const float qx = __shfl_sync(0xffffffffu, px, q), qy = __shfl_sync(0xffffffffu, py, q);
        const float qz = __shfl_sync(0xffffffffu, pz, q);
        const int qv = __shfl_sync(0xffffffffu, vrow, q);
        if (valid) *reinterpret_cast<float4*>(tile + q * NGLOD_KPAD + 4 * c) = sparse_gather4(sn, qx, qy, qz, qv, c);
    }
    __syncwarp();
    float d = 0.f;
    if (n_live) d = lane_decoder(sW, tile + lane * NGLOD_KPAD);
    __syncwarp();
    return d;
}

// Tensor-core path: gather straight into the group's A operand rows, then the shared MMA + epilogue.
__device__ __forceinline__ float tc_sparse_eval(const SparseDev& sn, TcGroup& g, float px, float py, float pz, int vrow,
                                                bool active) {
    const int lane = g.lane, row0 = g.wq * 32;
    const unsigned live = __ballot_sync(0xffffffffu, active);
    const int n_live = __popc(live);
    if (active) {
        g.idx[__popc(live & ((1u << lane) - 1u))] = lane;
        tc_store_split4(g.a_hi, g.a_lo, tc_elem_offset(row0 + lane, NGLOD_F), make_float4(px, py, pz, 1.f));
    }
    __syncwarp();
    const int sub = lane >> 3, c = lane & 7;
    for (int r = 0; r * 4 < n_live; ++r) {
        const int slot = r * 4 + sub;
        const bool valid = slot < n_live;
        const int q = g.idx[valid ? slot : 0];
        const float qx = __shfl_sync(0xffffffffu, px, q), qy = __shfl_sync(0xffffffffu, py, q);
        const float qz = __shfl_sync(0xffffffffu, pz, q);
        const int qv = __shfl_sync(0xffffffffu, vrow, q);
        if (valid) tc_store_split4(g.a_hi, g.a_lo, tc_elem_offset(row0 + q, 4 * c), sparse_gather4(sn, qx, qy, qz, qv, c));
    }
    __syncwarp();
    fence_proxy_async_smem();
    tc_fence_before_sync();
    named_bar_sync(g.bar_id, TCG_THREADS);
    if (g.wq == 0 && g.lane == 0) {
        tc_fence_after_sync();
        tc_issue_tile(g.tmem_acc, g.a_hi_s, g.a_lo_s, g.b_hi_s, g.b_lo_s);
        tc_commit(g.mbar_s);
    }
    mbar_wait(g.mbar_s, g.parity);
    g.parity ^= 1u;
    tc_fence_after_sync();
    return tc_epilogue(g.tmem_row, g.w1);
}

constexpr int SP_TC_GROUPS = 3;
constexpr int SP_TC_THREADS = SP_TC_GROUPS * TCG_THREADS;
constexpr int SP_TC_SMEM = TC_SMEM_BYTES(SP_TC_GROUPS);

struct EvalCtx {             // either path's per-thread context
    float* tile; int* idx; TcGroup grp; uint32_t tmem_base;
};

template <bool TC>
__device__ __forceinline__ void eval_prologue(const SparseDev& sn, char* smem_raw, EvalCtx& e) {
    if constexpr (TC) {
        e.tmem_base = tc_prologue(sn.dec, smem_raw, SP_TC_GROUPS);
        e.grp = tc_make_group(smem_raw, SP_TC_GROUPS, e.tmem_base);
    } else {
        float* smem = reinterpret_cast<float*>(smem_raw);
        sdf_stage_weights(sn.dec, smem);
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        e.tile = smem + SDF_SMEM_WARP_OFF + warp * SDF_SMEM_PER_WARP;
        e.idx = reinterpret_cast<int*>(e.tile + SDF_TILE_FLOATS);
        for (int k = lane; k < SDF_SMEM_PER_WARP; k += 32) e.tile[k] = 0.f;
        __syncthreads();
    }
}

template <bool TC>
__device__ __forceinline__ float eval_sparse(const SparseDev& sn, char* smem_raw, EvalCtx& e, float px, float py, float pz,
                                             int vrow, bool active) {
    if constexpr (TC) return tc_sparse_eval(sn, e.grp, px, py, pz, vrow, active);
    else return warp_sparse_eval(sn, reinterpret_cast<float*>(smem_raw), e.tile, e.idx, px, py, pz, vrow, active, threadIdx.x & 31);
}

template <bool TC>
__global__ void __launch_bounds__(TC ? SP_TC_THREADS : SDF_THREADS, TC ? 1 : 2)
sparse_sdf_forward_kernel(const SparseDev sn, const float* __restrict__ x, const int* __restrict__ pidx, const long long n,
                          float* __restrict__ out) {
    extern __shared__ __align__(128) char smem_raw[];
    EvalCtx e;
    eval_prologue<TC>(sn, smem_raw, e);
    const int lane = threadIdx.x & 31;
    constexpr int TILE = TC ? TCG_THREADS : 32;
    const long long unit = TC ? ((long long)blockIdx.x * SP_TC_GROUPS + (threadIdx.x >> 7)) : ((long long)blockIdx.x * SDF_WARPS + (threadIdx.x >> 5));
    const long long nunits = (long long)gridDim.x * (TC ? SP_TC_GROUPS : SDF_WARPS);
    for (long long base = unit * TILE; base < n; base += nunits * TILE) {
        const long long i = base + (TC ? ((threadIdx.x >> 5) & 3) * 32 : 0) + lane;
        // pidx < 0 = "no voxel holds this point" (SPC.query): such rows read no table and evaluate to 0
        const int pv = i < n ? __ldg(pidx + i) : -1;
        const bool active = pv >= 0;
        float px = 0.f, py = 0.f, pz = 0.f; int v = sn.vox_off;
        if (active) { px = __ldg(x + 3 * i); py = __ldg(x + 3 * i + 1); pz = __ldg(x + 3 * i + 2); v = sn.vox_off + pv; }
        const float d = eval_sparse<TC>(sn, smem_raw, e, px, py, pz, v, active);
        if (i < n) out[i] = active ? d : 0.f;
    }
    if constexpr (TC) tc_epilogue_free(e.tmem_base);
}

// ---- first voxel of a ray's nugget run from a query point (ray_aabb.cuh:42-192)
struct RayConst { float dx, dy, dz, ix, iy, iz, sx, sy, sz; };

__device__ __forceinline__ float aabb_voxel(const RayConst& rc, float qx, float qy, float qz, float vx, float vy, float vz, float r) {
    const float ox = qx - vx, oy = qy - vy, oz = qz - vz;
    const float cmax = fmaxf(fmaxf(fabsf(ox), fabsf(oy)), fabsf(oz));
    float winding = cmax < r ? -1.0f : 1.0f;
    winding *= r;
    if (winding < 0.f) return winding;
    const float d0 = __fmul_rn(__fmaf_rn(winding, rc.sx, -ox), rc.ix);
    const float d1 = __fmul_rn(__fmaf_rn(winding, rc.sy, -oy), rc.iy);
    const float d2 = __fmul_rn(__fmaf_rn(winding, rc.sz, -oz), rc.iz);
    const float ltxy = __fmaf_rn(rc.dy, d0, oy), ltxz = __fmaf_rn(rc.dz, d0, oz);
    const float ltyx = __fmaf_rn(rc.dx, d1, ox), ltyz = __fmaf_rn(rc.dz, d1, oz);
    const float ltzx = __fmaf_rn(rc.dx, d2, ox), ltzy = __fmaf_rn(rc.dy, d2, oy);
    if ((d0 >= 0.0f) && (fabsf(ltxy) < r) && (fabsf(ltxz) < r)) return d0;
    if ((d1 >= 0.0f) && (fabsf(ltyx) < r) && (fabsf(ltyz) < r)) return d1;
    if ((d2 >= 0.0f) && (fabsf(ltzx) < r) && (fabsf(ltzy) < r)) return d2;
    return 0.0f;
}

struct SpTraceParams { int num_steps; int compute_normals; float min_dis; float osc; float far; float h; };

enum : int { SP_EMPTY = 0, SP_MARCH = 1, SP_N0 = 2 };     // SP_N0..SP_N0+5: normal taps

template <bool TC>
__global__ void __launch_bounds__(TC ? SP_TC_THREADS : SDF_THREADS, TC ? 1 : 2)
spc_sphere_trace_kernel(const SparseDev sn, const int2* __restrict__ nuggets, const int* __restrict__ run_begin,
                        const int* __restrict__ run_end, const float* __restrict__ ray_o, const float* __restrict__ ray_d, const long long n,
                        const SpTraceParams tp, float* __restrict__ out_x, float* __restrict__ out_t,
                        uint8_t* __restrict__ out_hit, float* __restrict__ out_n, int* __restrict__ out_pidx,
                        int* __restrict__ queue, unsigned long long* __restrict__ stats) {
    extern __shared__ __align__(128) char smem_raw[];
    EvalCtx e;
    eval_prologue<TC>(sn, smem_raw, e);
    const int lane = threadIdx.x & 31;
    const float vr = 1.0f / (float)(1 << (sn.lod + sn.base_lod));       // voxel "radius" in [-1,1] units

    int phase = SP_EMPTY, iter = 0, pidx = -1, beg = 0, end = 0, cur = 0;
    long long ray = -1;
    float ox = 0.f, oy = 0.f, oz = 0.f;
    RayConst rc = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float x = 0.f, y = 0.f, z = 0.f, t = 0.f, dprev = 0.f, gtmp = 0.f, g0 = 0.f, g1 = 0.f, g2 = 0.f;
    bool exhausted = false;
    unsigned long long n_eval = 0;
    const unsigned lt_mask = (1u << lane) - 1u;

    auto retire = [&](bool hit, float nx, float ny, float nz) {
        out_x[3 * ray] = x; out_x[3 * ray + 1] = y; out_x[3 * ray + 2] = z;
        out_t[ray] = t;
        out_hit[ray] = hit ? 1 : 0;
        out_n[3 * ray] = nx; out_n[3 * ray + 1] = ny; out_n[3 * ray + 2] = nz;
        if (out_pidx) out_pidx[ray] = pidx;
        phase = SP_EMPTY;
    };
    // walk the run front to back from (x,y,z); returns false (and parks the ray at t = 100) when no voxel is left.
    // The reference walks the whole run from its first nugget at every step (ray_aabb.cuh:104-192).  The run is ordered
    // along the ray, so after a FORWARD step every nugget in front of the voxel the point was in lies behind the point and
    // answers 0 ("not inside, entry plane behind"): the walk may resume at that voxel (`cur`) with the same result -- at
    // level 7 a run that crosses the surface shell is ~15 nuggets of two dependent loads each, per march step.  A backward
    // step (d < 0, the point is inside the surface) restarts at the first nugget.
    auto locate = [&](bool resume) -> bool {
        for (int i = resume ? cur : beg; i < end; ++i) {
            const int pi = nuggets[i].y;
            const short4 p = __ldg(sn.voxels + sn.vox_off + pi);
            const float vx = __fmaf_rn(vr, __fmaf_rn(2.0f, (float)p.x, 1.0f), -1.0f);
            const float vy = __fmaf_rn(vr, __fmaf_rn(2.0f, (float)p.y, 1.0f), -1.0f);
            const float vz = __fmaf_rn(vr, __fmaf_rn(2.0f, (float)p.z, 1.0f), -1.0f);
            const float d = aabb_voxel(rc, x, y, z, vx, vy, vz, vr);
            if (d != 0.0f) {
                pidx = pi;
                cur = i;
                if (d > 0.0f) {
                    t = t + d;
                    x = __fmaf_rn(rc.dx, t, ox); y = __fmaf_rn(rc.dy, t, oy); z = __fmaf_rn(rc.dz, t, oz);
                }
                return true;
            }
        }
        t = 100.0f;
        x = __fmaf_rn(rc.dx, t, ox); y = __fmaf_rn(rc.dy, t, oy); z = __fmaf_rn(rc.dz, t, oz);
        return false;
    };

    for (;;) {
#pragma unroll 1
        for (int attempt = 0; attempt < NGLOD_SPC_REFILL_ATTEMPTS && !exhausted; ++attempt) {
            const unsigned free_mask = __ballot_sync(0xffffffffu, phase == SP_EMPTY);
            if (!free_mask) break;
            const int nfree = __popc(free_mask);
#if NGLOD_SPC_REFILL_MIN > 1
            if (nfree < NGLOD_SPC_REFILL_MIN) break;
#endif
            int base = 0;
            if (lane == 0) base = atomicAdd(queue, nfree);
            base = __shfl_sync(0xffffffffu, base, 0);
            if ((long long)base + nfree >= n) exhausted = true;
            if (phase == SP_EMPTY) {
                const long long i = (long long)base + __popc(free_mask & lt_mask);
                if (i < n) {
                    ray = i;
                    ox = __ldg(ray_o + 3 * i); oy = __ldg(ray_o + 3 * i + 1); oz = __ldg(ray_o + 3 * i + 2);
                    rc.dx = __ldg(ray_d + 3 * i); rc.dy = __ldg(ray_d + 3 * i + 1); rc.dz = __ldg(ray_d + 3 * i + 2);
                    rc.ix = 1.0f / rc.dx; rc.iy = 1.0f / rc.dy; rc.iz = 1.0f / rc.dz;
                    rc.sx = signbit(rc.dx) ? 1.0f : -1.0f; rc.sy = signbit(rc.dy) ? 1.0f : -1.0f; rc.sz = signbit(rc.dz) ? 1.0f : -1.0f;
                    beg = run_begin[i]; end = run_end[i];
                    x = ox; y = oy; z = oz; t = 0.f; dprev = 0.f; iter = 0; pidx = -1; cur = beg;
                    if (beg == end) retire(false, 0.f, 0.f, 0.f);                 // never entered the octree
                    else if (tp.num_steps > 0 && locate(false)) phase = SP_MARCH;
                    else retire(false, 0.f, 0.f, 0.f);
                }
            }
        }
        const bool occupied = phase != SP_EMPTY;
        const unsigned act = __ballot_sync(0xffffffffu, occupied);
        if constexpr (TC) {
            if (!tc_group_any(e.grp.bar_id, occupied || !exhausted)) break;
        } else {
            if (!act) { if (exhausted) break; continue; }
        }
        float qx = x, qy = y, qz = z;
        if (phase >= SP_N0) {
            const int m = phase - SP_N0;
            const float eps = (m & 1) ? -tp.h : tp.h;
            const int axis = m >> 1;
            if (axis == 0) qx = x + eps; else if (axis == 1) qy = y + eps; else qz = z + eps;
        }
        const float dv = eval_sparse<TC>(sn, smem_raw, e, qx, qy, qz, sn.vox_off + max(pidx, 0), occupied);
        if (lane == 0) n_eval += __popc(act);
        if (phase == SP_MARCH) {
            // step.cuh:31-86
            const float d = dv;
            t = t + d;
            bool hit = (double)fabsf(d) < (double)tp.min_dis;
            hit |= ((double)fabsf(d + dprev) * 0.5) < (double)tp.osc;
            bool cond = (t < tp.far) && !hit;
            x = __fmaf_rn(rc.dx, t, ox); y = __fmaf_rn(rc.dy, t, oy); z = __fmaf_rn(rc.dz, t, oz);
            dprev = d;
            ++iter;
            if (cond) cond = locate(d >= 0.f);              // SDF.cu:442-460: re-locate the voxel from the new x
            if (hit) {
                if (tp.compute_normals) phase = SP_N0; else retire(true, 0.f, 0.f, 0.f);
            } else if (!cond || iter >= tp.num_steps) {
                retire(false, 0.f, 0.f, 0.f);
            }
        } else if (phase >= SP_N0) {
            const int m = phase - SP_N0;
            if ((m & 1) == 0) { gtmp = dv; phase = phase + 1; }
            else {
                const float g = gtmp - dv;
                if (m == 1) g0 = g; else if (m == 3) g1 = g; else g2 = g;
                if (m == 5) {
                    const float nrm = sqrtf(g0 * g0 + g1 * g1 + g2 * g2);
                    const float inv = nrm > 0.f ? 1.0f / nrm : 0.f;
                    retire(true, g0 * inv, g1 * inv, g2 * inv);
                } else phase = phase + 1;
            }
        }
    }
    if (stats && lane == 0) atomicAdd(stats, n_eval);
    if constexpr (TC) tc_epilogue_free(e.tmem_base);
}

}  // namespace

extern "C" int nglod_sparse_sdf_forward(const nglod_sparse_net_t* net, int32_t lod, const float* x, const int32_t* pidx,
                                        int64_t n, float* out, void* stream) {
    SparseDev sn;
    if (int e = make_sparse_dev(net, lod, sn)) return e;
    if (n < 0 || (n > 0 && (!x || !pidx || !out))) return NGLOD_EINVAL;
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (net->math_mode == NGLOD_MATH_TC3XTF32) {
        auto kern = sparse_sdf_forward_kernel<true>;
        NGLOD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SP_TC_SMEM));
        long long grid = nglod_sm_count();
        const long long want = (n + SP_TC_THREADS - 1) / SP_TC_THREADS;
        if (want < grid) grid = want;
        kern<<<(int)grid, SP_TC_THREADS, SP_TC_SMEM, st>>>(sn, x, pidx, (long long)n, out);
    } else {
        auto kern = sparse_sdf_forward_kernel<false>;
        NGLOD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SDF_SMEM_BYTES));
        long long grid = (long long)nglod_sm_count() * 2;
        const long long want = (n + SDF_THREADS - 1) / SDF_THREADS;
        if (want < grid) grid = want;
        kern<<<(int)grid, SDF_THREADS, SDF_SMEM_BYTES, st>>>(sn, x, pidx, (long long)n, out);
    }
    return (int)cudaGetLastError();
}

extern "C" int nglod_spc_sphere_trace(const nglod_sparse_net_t* net, int32_t lod, const int32_t* nuggets,
                                      const int32_t* offsets, const float* ray_o, const float* ray_d, int64_t n,
                                      const nglod_trace_opts_t* opts, float* x, float* depth, uint8_t* hit, float* normal,
                                      int32_t* pidx_out, int32_t* queue, unsigned long long* stats, void* stream) {
    if (n > 0 && !offsets) return NGLOD_EINVAL;
    return nglod_spc_sphere_trace_runs(net, lod, nuggets, offsets, offsets ? offsets + 1 : nullptr, ray_o, ray_d, n, opts, x,
                                       depth, hit, normal, pidx_out, queue, stats, stream);
}

extern "C" int nglod_spc_sphere_trace_runs(const nglod_sparse_net_t* net, int32_t lod, const int32_t* nuggets,
                                           const int32_t* run_begin, const int32_t* run_end, const float* ray_o,
                                           const float* ray_d, int64_t n, const nglod_trace_opts_t* opts, float* x,
                                           float* depth, uint8_t* hit, float* normal, int32_t* pidx_out, int32_t* queue,
                                           unsigned long long* stats, void* stream) {
    SparseDev sn;
    if (int e = make_sparse_dev(net, lod, sn)) return e;
    if (!opts || n < 0 || n > 2000000000ll || opts->num_steps < 0) return NGLOD_EINVAL;
    if (n == 0) return 0;
    if (!run_begin || !run_end || !ray_o || !ray_d || !x || !depth || !hit || !normal || !queue) return NGLOD_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    NGLOD_CUDA_TRY(cudaMemsetAsync(queue, 0, sizeof(int32_t), st));
    SpTraceParams tp;
    tp.num_steps = opts->num_steps;
    tp.compute_normals = opts->compute_normals;
    tp.min_dis = (float)opts->min_dis;
    tp.osc = (float)(opts->min_dis * 5.0);
    tp.far = (float)opts->far;
    tp.h = (float)opts->normal_h;
    const int2* nug = reinterpret_cast<const int2*>(nuggets);
    if (net->math_mode == NGLOD_MATH_TC3XTF32) {
        auto kern = spc_sphere_trace_kernel<true>;
        NGLOD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SP_TC_SMEM));
        long long grid = nglod_sm_count();
        const long long want = (n + SP_TC_THREADS - 1) / SP_TC_THREADS;
        if (want < grid) grid = want;
        kern<<<(int)grid, SP_TC_THREADS, SP_TC_SMEM, st>>>(sn, nug, run_begin, run_end, ray_o, ray_d, (long long)n, tp, x, depth, hit,
                                                           normal, pidx_out, queue, stats);
    } else {
        auto kern = spc_sphere_trace_kernel<false>;
        NGLOD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SDF_SMEM_BYTES));
        long long grid = (long long)nglod_sm_count() * 2;
        const long long want = (n + SDF_THREADS - 1) / SDF_THREADS;
        if (want < grid) grid = want;
        kern<<<(int)grid, SDF_THREADS, SDF_SMEM_BYTES, st>>>(sn, nug, run_begin, run_end, ray_o, ray_d, (long long)n, tp, x, depth, hit,
                                                             normal, pidx_out, queue, stats);
    }
    return (int)cudaGetLastError();
}
