// nglod_b200 -- persistent-thread sphere tracer with the SDF evaluated inline.
//
// Behavioural spec: SphereTracer.forward, sdf-net/lib/tracer/SphereTracer.py:41-132,
// restated as an independent per-ray state machine (the batch loop there has no
// cross-ray coupling except a global early-out, which cannot change any ray's
// result):
//
//   (x, t, live) = aabb(o, dir)                                   :53
//   live:  d = dprev = sdf(x)                  [INIT]             :64-66
//   for i in 0..num_steps-1:                                       :74
//       flag  = |t| < far                                          :84
//       live &= |d| > min_dis  &  |(d+dprev)/2| > 3*min_dis & flag :87-93
//       live:  x = o + dir*t ; dprev = d                           :102-105
//              d = sdf(x)*step_size ; t += d   [MARCH]             :109-114
//   hit = flag & all(|x| <= 1)                                     :119
//   hit:   normal = normalize(finitediff(x), eps=1e-5)  [N0..N5]   :128-130
//
// One warp = 32 ray slots.  Every round each occupied slot contributes one
// query point (march position or one of the six normal taps); the warp
// evaluates them cooperatively (sdf_core.cuh), each lane consumes its value
// and advances its state machine.  Finished slots are refilled from a global
// atomic queue (ballot + popc ranks), so lanes stay occupied until the frame
// runs dry: no host round trip, no per-step launches, no cond.any() sync.
#include "sdf_core.cuh"
#include "sdf_tc.cuh"
#include "aabb.cuh"

namespace {

enum : int { PH_EMPTY = 0, PH_INIT = 1, PH_MARCH = 2, PH_N0 = 3 };   // PH_N0..PH_N0+5: normal taps

struct TraceParams {
    int num_steps;
    int compute_normals;
    float step_size, min_dis, min_dis3, far, h, two_h;
};

#ifndef NGLOD_TRACE_GROUPS
#define NGLOD_TRACE_GROUPS 3
#endif
#ifndef NGLOD_TRACE_GROUPS_SINGLE
#define NGLOD_TRACE_GROUPS_SINGLE 4
#endif
// how many times a warp goes back to the queue in one round while it still has empty lanes (rays that miss the box retire on
// the spot and free their lane again).  Measured on the bench frame (profiles/exp_attempts.sh), device frame / host-to-host
// frame in ms: 1: 0.864 / 1.048, 2: 0.877 / 1.074, 4: 0.872 / 1.079, 8: 0.886 / 1.096, 32: 0.898 / 1.114 -- draining the
// box-missing rays faster only lengthens the rounds of the rays that march (and bunches their stores).
#ifndef NGLOD_TRACE_REFILL_ATTEMPTS
#define NGLOD_TRACE_REFILL_ATTEMPTS 1
#endif
// a warp goes back to the queue only once it has this many empty lanes: the atomic's round trip sits on the round's critical
// path, so it is paid per batch of rays rather than per retired ray.  Same frame (profiles/exp_refill_min.sh), device / host-to-
// host (L2 flushed) in ms: 1: 0.864 / 1.060, 4: 0.854 / 1.038, 8: 0.838 / 1.035, 16: 0.858 / 1.040.  (6 and 12 on a second box: 0.850 / 0.848.)  Results do not depend on it
// (a ray's march never looks at its lane or its neighbours).
#ifndef NGLOD_TRACE_REFILL_MIN
#define NGLOD_TRACE_REFILL_MIN 8
#endif
constexpr int trace_groups(int mode) { return mode == TC_MULTI ? NGLOD_TRACE_GROUPS : NGLOD_TRACE_GROUPS_SINGLE; }
constexpr int trace_tc_threads(int mode) { return trace_groups(mode) * TCG_THREADS; }
constexpr int trace_tc_smem(int mode) { return TC_SMEM_BYTES_W(trace_groups(mode), tc_mode_scratch(mode)); }

// TC = false: FP32 CUDA-core decoder, warps are independent (8 per CTA, 2 CTAs/SM).
// TC = true : tcgen05 decoder; 4 warps form a 128-row MMA tile and advance in lock-step rounds.  MODE picks the
//             gather (sdf_tc.cuh): per-LOD fp32 grids, or ONE prefix-summed grid in fp32 / fp16 x-pair lines.
template <bool TC, int MODE>
__global__ void __launch_bounds__(TC ? trace_tc_threads(MODE) : SDF_THREADS, TC ? 1 : 2)
sphere_trace_kernel(const NetDev net, const float* __restrict__ ray_o, const float* __restrict__ ray_d,
                    const long long n, const TraceParams tp, float* __restrict__ out_x,
                    float* __restrict__ out_t, uint8_t* __restrict__ out_hit, float* __restrict__ out_n,
                    int* __restrict__ queue, unsigned long long* __restrict__ stats
                    ) {
    extern __shared__ __align__(128) char smem_raw[];
    float* smem = reinterpret_cast<float*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* tile = nullptr; int* idx = nullptr;
    uint32_t tmem_base = 0;
    TcGroup grp;
    if constexpr (TC) {
        tmem_base = tc_prologue(net, smem_raw, trace_groups(MODE), tc_mode_scratch(MODE));
        grp = tc_make_group(smem_raw, trace_groups(MODE), tmem_base, tc_mode_scratch(MODE));
    } else {
        sdf_stage_weights(net, smem);
        tile = smem + SDF_SMEM_WARP_OFF + warp * SDF_SMEM_PER_WARP;
        idx = reinterpret_cast<int*>(tile + SDF_TILE_FLOATS);
        for (int e = lane; e < SDF_SMEM_PER_WARP; e += 32) tile[e] = 0.f;
        __syncthreads();
    }

    // per-lane ray slot
    int phase = PH_EMPTY, step = 0;
    long long ray = -1;
    float ox = 0.f, oy = 0.f, oz = 0.f, dx = 0.f, dy = 0.f, dz = 0.f;
    float x = 0.f, y = 0.f, z = 0.f, t = 0.f, d = 0.f, dprev = 0.f;
    float g0 = 0.f, g1 = 0.f, g2 = 0.f, gtmp = 0.f;
    bool flag = false;
    bool exhausted = false;                     // warp-uniform: the queue has run past n
    unsigned long long n_eval = 0, n_march = 0; // lane 0 only
    const unsigned lt_mask = (1u << lane) - 1u;

    auto retire = [&](bool hit, float nx, float ny, float nz) {
        if (out_t == nullptr) {
            // packed records (nglod_sphere_trace_packed) -- what a pinned HOST buffer wants: posted PCIe writes as the rays
            // retire (the eight scalar stores of the field layout cost the kernel 0.87 -> 1.33 ms when they cross PCIe).
            // out_hit == nullptr: out_x is [n][8] = {depth, nx, ny, nz | hit (u32 0/1), x, y, z}, two 16-byte stores per ray;
            // otherwise out_x is [n][4] = {depth, nx, ny, nz}, one 16-byte store, and the hit byte goes to out_hit
            if (out_hit == nullptr) {
                float4* rec = reinterpret_cast<float4*>(out_x) + 2 * ray;
                rec[0] = make_float4(t, nx, ny, nz);
                rec[1] = make_float4(__uint_as_float(hit ? 1u : 0u), x, y, z);
            } else {
                reinterpret_cast<float4*>(out_x)[ray] = make_float4(t, nx, ny, nz);
                out_hit[ray] = hit ? 1 : 0;
            }
        } else {
            out_x[3 * ray] = x; out_x[3 * ray + 1] = y; out_x[3 * ray + 2] = z;
            out_t[ray] = t;
            out_hit[ray] = hit ? 1 : 0;
            out_n[3 * ray] = nx; out_n[3 * ray + 1] = ny; out_n[3 * ray + 2] = nz;
        }
        phase = PH_EMPTY;
    };
    auto finish_march = [&]() {
        const bool outside = (fabsf(x) > 1.0f) || (fabsf(y) > 1.0f) || (fabsf(z) > 1.0f);
        const bool hit = flag && !outside;
        if (hit && tp.compute_normals) phase = PH_N0;
        else retire(hit, 0.f, 0.f, 0.f);
    };
    auto march_check = [&](bool live) {
        if (step < tp.num_steps) {
            flag = fabsf(t) < tp.far;
            live = live && (fabsf(d) > tp.min_dis) && (fabsf((d + dprev) * 0.5f) > tp.min_dis3) && flag;
            if (live) {
                // torch.addcmul(ray_o, ray_d, t): product rounded, then added (no fma)
                x = __fadd_rn(ox, __fmul_rn(dx, t));
                y = __fadd_rn(oy, __fmul_rn(dy, t));
                z = __fadd_rn(oz, __fmul_rn(dz, t));
                dprev = d;
                phase = PH_MARCH;
                return;
            }
        }
        finish_march();
    };

    // this round's query point of the lane's ray: its march position or one of the six normal taps
    auto query_point = [&](float& qx, float& qy, float& qz) {
        qx = x; qy = y; qz = z;
        if (phase >= PH_N0) {
            const int m = phase - PH_N0;
            const float e = (m & 1) ? -tp.h : tp.h;
            const int axis = m >> 1;
            if (axis == 0) qx = x + e; else if (axis == 1) qy = y + e; else qz = z + e;
        }
    };
    // consume the value of that query and advance the lane's state machine
    auto advance = [&](float dv) {
        if (phase == PH_INIT) {
            d = dv; dprev = dv; step = 0;
            march_check(true);
        } else if (phase == PH_MARCH) {
            d = dv * tp.step_size;
            t = t + d;
            ++step;
            march_check(true);
        } else if (phase >= PH_N0) {
            const int m = phase - PH_N0;
            if ((m & 1) == 0) {
                gtmp = dv;
                phase = phase + 1;
            } else {
                const float g = (gtmp - dv) / tp.two_h;
                if (m == 1) g0 = g; else if (m == 3) g1 = g; else g2 = g;
                if (m == 5) {
                    // F.normalize(grad, p=2, dim=-1, eps=1e-5)
                    const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(g0, g0), __fmul_rn(g1, g1)), __fmul_rn(g2, g2)));
                    const float den = fmaxf(nrm, 1e-5f);
                    retire(true, g0 / den, g1 / den, g2 / den);
                } else {
                    phase = phase + 1;
                }
            }
        }
    };

#ifdef NGLOD_TRACE_TIMING
    if constexpr (TC) { for (int i = 0; i < 8; ++i) grp.tm[i] = 0; grp.t_last = clock64(); }
#define TR_TICK(i) do { if constexpr (TC) { TcGroup& g = grp; TC_TICK(i); } } while (0)
#else
#define TR_TICK(i) do { } while (0)
#endif
    // refill empty slots from the global queue (ballot + popc ranks; rays that miss the box retire on the spot)
    auto refill = [&]() {
#pragma unroll 1
        for (int attempt = 0; attempt < NGLOD_TRACE_REFILL_ATTEMPTS && !exhausted; ++attempt) {
            const unsigned free_mask = __ballot_sync(0xffffffffu, phase == PH_EMPTY);
            if (!free_mask) break;
            const int nfree = __popc(free_mask);
#if NGLOD_TRACE_REFILL_MIN > 1
            if (nfree < NGLOD_TRACE_REFILL_MIN) break;     // the atomic's round trip is on the round's critical path: batch it
#endif
            int base = 0;
            if (lane == 0) base = atomicAdd(queue, nfree);
            base = __shfl_sync(0xffffffffu, base, 0);
            if ((long long)base + nfree >= n) exhausted = true;
            if (phase == PH_EMPTY) {
                const long long i = (long long)base + __popc(free_mask & lt_mask);
                if (i < n) {
                    ray = i;
                    ox = __ldg(ray_o + 3 * i); oy = __ldg(ray_o + 3 * i + 1); oz = __ldg(ray_o + 3 * i + 2);
                    dx = __ldg(ray_d + 3 * i); dy = __ldg(ray_d + 3 * i + 1); dz = __ldg(ray_d + 3 * i + 2);
                    const AabbResult r = ray_unit_cube(ox, oy, oz, dx, dy, dz);
                    x = r.x; y = r.y; z = r.z; t = r.t;
                    step = 0; flag = false; d = 0.f; dprev = 0.f;
                    if (r.hit) phase = PH_INIT;
                    else march_check(false);      // never marched: retire now or go take normals
                }
            }
        }
    };
    for (;;) {
        refill();
        TR_TICK(0);
        const bool occupied = phase != PH_EMPTY;
        const unsigned act = __ballot_sync(0xffffffffu, occupied);
        if constexpr (!TC) {
            if (!act) {
                if (exhausted) break;
                continue;
            }
            float qx, qy, qz;
            query_point(qx, qy, qz);
            const float dv = warp_sdf_eval(net, smem, tile, idx, qx, qy, qz, occupied, lane);
            const unsigned march_mask = __ballot_sync(0xffffffffu, phase == PH_MARCH);
            if (lane == 0) { n_eval += __popc(act); n_march += __popc(march_mask); }
            advance(dv);
        } else {
            float qx, qy, qz;
            query_point(qx, qy, qz);
            const unsigned march_mask = __ballot_sync(0xffffffffu, phase == PH_MARCH);
            if (lane == 0) { n_eval += __popc(act); n_march += __popc(march_mask); }
            // the 4 warps of a tile leave together: the vote (a slot still occupied, or a queue not yet dry) rides on
            // the barrier that precedes the MMA, so a round costs one group barrier
            float dv = 0.f;
            if (!tc_group_eval_any<MODE>(net, grp, qx, qy, qz, occupied, occupied || !exhausted, dv)) break;
            advance(dv);
        }
        TR_TICK(5);
    }
    if (stats && lane == 0) {
        atomicAdd(stats, n_eval);
        atomicAdd(stats + 1, n_march);
#ifdef NGLOD_TRACE_TIMING
        if constexpr (TC) for (int i = 0; i < 6; ++i) atomicAdd(stats + 2 + i, (unsigned long long)grp.tm[i]);
#endif
    }
    if constexpr (TC) tc_epilogue_free(tmem_base);
}

}  // namespace

static int trace_launch(const nglod_net_t* net, int32_t lod, const float* ray_o, const float* ray_d, int64_t n,
                        const nglod_trace_opts_t* opts, float* x, float* depth, uint8_t* hit, float* normal, int32_t* queue,
                        unsigned long long* stats, void* stream);

extern "C" int nglod_sphere_trace(const nglod_net_t* net, int32_t lod, const float* ray_o, const float* ray_d,
                                  int64_t n, const nglod_trace_opts_t* opts, float* x, float* depth,
                                  uint8_t* hit, float* normal, int32_t* queue, unsigned long long* stats,
                                  void* stream) {
    if (n > 0 && (!x || !depth || !hit || !normal)) return NGLOD_EINVAL;
    return trace_launch(net, lod, ray_o, ray_d, n, opts, x, depth, hit, normal, queue, stats, stream);
}

extern "C" int nglod_sphere_trace_packed(const nglod_net_t* net, int32_t lod, const float* ray_o, const float* ray_d,
                                         int64_t n, const nglod_trace_opts_t* opts, float* packed, uint8_t* hit,
                                         int32_t* queue, unsigned long long* stats, void* stream) {
    if (n > 0 && (!packed || (reinterpret_cast<uintptr_t>(packed) & (hit ? 15u : 31u)))) return NGLOD_EINVAL;
    return trace_launch(net, lod, ray_o, ray_d, n, opts, packed, nullptr, hit, nullptr, queue, stats, stream);
}

extern "C" int nglod_sphere_trace_camera(const nglod_net_t* net, int32_t lod, const float* origin, const float* view,
                                         const float* right, const float* up, float tan_half_fov, int32_t ortho,
                                         const float* window_x, const float* window_y, int32_t width, int32_t height,
                                         const nglod_trace_opts_t* opts, float* workspace, float* packed, uint8_t* hit,
                                         uint8_t* hit_copy, int32_t* queue, unsigned long long* stats, void* stream) {
    if (width < 0 || height < 0) return NGLOD_EINVAL;
    const long long n = (long long)width * height;
    if (n == 0) return 0;
    if (!workspace || !window_x || !window_y || !packed || (hit_copy && !hit)) return NGLOD_EINVAL;
    if (reinterpret_cast<uintptr_t>(packed) & (hit ? 15u : 31u)) return NGLOD_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    float* ray_o = workspace;
    float* ray_d = workspace + 3 * n;
    float* wx = workspace + 6 * n;
    float* wy = wx + width;
    // the window may live in (pinned) host memory or on the device: 4 (W + H) bytes
    NGLOD_CUDA_TRY(cudaMemcpyAsync(wx, window_x, sizeof(float) * width, cudaMemcpyDefault, st));
    NGLOD_CUDA_TRY(cudaMemcpyAsync(wy, window_y, sizeof(float) * height, cudaMemcpyDefault, st));
    if (int e = nglod_generate_rays(origin, view, right, up, tan_half_fov, ortho, wx, wy, width, height, ray_o, ray_d, stream))
        return e;
    if (int e = trace_launch(net, lod, ray_o, ray_d, n, opts, packed, nullptr, hit, nullptr, queue, stats, stream)) return e;
    if (hit_copy) NGLOD_CUDA_TRY(cudaMemcpyAsync(hit_copy, hit, (size_t)n, cudaMemcpyDefault, st));
    return 0;
}

// depth == nullptr: x is the packed record buffer (32-byte records if hit == nullptr too, else 16-byte records + hit bytes)
static int trace_launch(const nglod_net_t* net, int32_t lod, const float* ray_o, const float* ray_d, int64_t n,
                        const nglod_trace_opts_t* opts, float* x, float* depth, uint8_t* hit, float* normal, int32_t* queue,
                        unsigned long long* stats, void* stream) {
    if (int e = nglod_check_net(net, lod)) return e;
    if (!opts || n < 0 || n > 2000000000ll) return NGLOD_EINVAL;
    if (n == 0) return 0;
    if (!ray_o || !ray_d || !x || !queue) return NGLOD_EINVAL;
    if (opts->num_steps < 0 || opts->max_ctas < 0) return NGLOD_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    NGLOD_CUDA_TRY(cudaMemsetAsync(queue, 0, sizeof(int32_t), st));
    TraceParams tp;
    tp.num_steps = opts->num_steps;
    tp.compute_normals = opts->compute_normals;
    tp.step_size = (float)opts->step_size;
    tp.min_dis = (float)opts->min_dis;
    tp.min_dis3 = (float)(opts->min_dis * 3.0);      // torch: tensor > (python float * 3) -> float32(min_dis*3)
    tp.far = (float)opts->far;
    tp.h = (float)opts->normal_h;
    tp.two_h = (float)(opts->normal_h * 2.0);
    const NetDev nd = nglod_make_netdev_infer(net, lod);
    if (net->math_mode == NGLOD_MATH_TC3XTF32) {
        const int mode = nd.half_pairs ? TC_SINGLE_HALF : (nd.num_lods == 1 ? TC_SINGLE_F32 : TC_MULTI);
        auto kern = mode == TC_SINGLE_HALF ? sphere_trace_kernel<true, TC_SINGLE_HALF>
                  : mode == TC_SINGLE_F32 ? sphere_trace_kernel<true, TC_SINGLE_F32> : sphere_trace_kernel<true, TC_MULTI>;
        const int threads = trace_tc_threads(mode), smem = trace_tc_smem(mode);
        NGLOD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        long long grid = nglod_sm_count();
        const long long want = (n + threads - 1) / threads;
        if (want < grid) grid = want;
        if (opts->max_ctas > 0 && opts->max_ctas < grid) grid = opts->max_ctas;
        kern<<<(int)grid, threads, smem, st>>>(nd, ray_o, ray_d, (long long)n, tp, x, depth, hit, normal, queue, stats);
        return (int)cudaGetLastError();
    }
    auto kern = sphere_trace_kernel<false, TC_MULTI>;
    NGLOD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SDF_SMEM_BYTES));
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SDF_THREADS, SDF_SMEM_BYTES) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    long long grid = (long long)nglod_sm_count() * per_sm;
    const long long want = (n + SDF_THREADS - 1) / SDF_THREADS;
    if (want < grid) grid = want;
    if (opts->max_ctas > 0 && (long long)opts->max_ctas * per_sm < grid) grid = (long long)opts->max_ctas * per_sm;
    kern<<<(int)grid, SDF_THREADS, SDF_SMEM_BYTES, st>>>(nd, ray_o, ray_d, (long long)n, tp, x, depth, hit, normal,
                                                         queue, stats
                                                         );
    return (int)cudaGetLastError();
}
