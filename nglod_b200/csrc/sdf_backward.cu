// nglod_b200 -- backward of OctreeSDF.sdf(x, lod) and the fused L2 training step.
//
// Reference behaviour: autograd through sdf-net/lib/models/OctreeSDF.py:94-146
// (grid_sampler_3d_backward scatter + Linear grads), loss of
// sdf-net/lib/trainer.py:317-339.
//
// One kernel, no saved activations (the forward is recomputed in-kernel, which
// costs ~6k FMA/query but removes the ~0.8 kB/query of autograd state the
// reference round-trips through HBM).  Per warp batch of 32 queries:
//   A  gather: tile[q] = {feat, xyz, 1}                      (sdf_core.cuh)
//   B  thread-per-query: pre[j] -> smem P[q][j]; d; g_d (given, or 2*(d-gt)*scale);
//      g_in[k] = sum_j W0[j][k] * g_h[j]  with g_h[j] = g_d*W1[j]*[pre_j>0]
//   C  lane owns hidden units {lane, lane+32, lane+64, lane+96}: 144 register
//      accumulators of dW0 (incl. db0 as column 35), dW1, db1 summed over the
//      warp's queries -- flushed ONCE per CTA at kernel end (smem, then RED)
//   D  scatter: 8 lanes per corner, one red.global.add.v4.f32 per lane per
//      corner (a full 128-byte line per corner per query), for every LOD <= lod;
//      optional dL/dx with PyTorch's border-clip rule.
#include "sdf_core.cuh"
#include "sparse_core.cuh"
#include "internal.h"
#include <cstdlib>

namespace {

#define BWD_P_STRIDE (NGLOD_H + 1)                        // padded: lane q writes column j conflict-free
#define BWD_P_FLOATS (32 * BWD_P_STRIDE)
#define BWD_PER_WARP (SDF_SMEM_PER_WARP + BWD_P_FLOATS + 32)  // tile, idx, P, gd
#define BWD_W2_OFF (SDF_SMEM_WARP_OFF + SDF_WARPS * BWD_PER_WARP)   // W0|b0 again, hidden units interleaved in pairs:
                                                                   // W2[j/2][k] = {W[j][k], W[j+1][k]} (for FFMA2 over j)
#define BWD_SMEM_BYTES ((BWD_W2_OFF + SDF_W0_FLOATS) * 4)
#define BWD_ACC_FLOATS (SDF_W0_FLOATS + NGLOD_H + 4)       // dW0|db0 (H x 36), dW1 (H), db1

struct BwdAxis {
    int i0, i1;
    float w0, w1;
    float mult;     // d(index)/d(p) : R/2 inside, 0 where grid_sample clips (u<=0 or u>=R)
};

__device__ __forceinline__ BwdAxis bwd_axis(float p, int R) {
    BwdAxis a;
    const float fR = (float)R;
    const float u_raw = ((p + 1.f) * 0.5f) * fR;
    a.mult = (u_raw > 0.f && u_raw < fR) ? fR * 0.5f : 0.f;
    const float u = fminf(fR, fmaxf(u_raw, 0.f));
    const float f0 = floorf(u);
    a.i0 = (int)f0;
    a.i1 = min(a.i0 + 1, R);
    a.w0 = (f0 + 1.f) - u;
    a.w1 = u - f0;
    return a;
}

template <bool FUSED_LOSS, bool WITH_GX>
__global__ void __launch_bounds__(SDF_THREADS, 1)
sdf_backward_kernel(const NetDev net, const GradDev grad, const float* __restrict__ x, const long long n,
                    const float* __restrict__ grad_out, const float* __restrict__ gt, const float loss_scale,
                    float* __restrict__ grad_x, float* __restrict__ loss_out) {
    extern __shared__ __align__(16) float smem[];
    sdf_stage_weights(net, smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* wbase = smem + SDF_SMEM_WARP_OFF + warp * BWD_PER_WARP;
    float* tile = wbase;
    int* idx = reinterpret_cast<int*>(wbase + SDF_TILE_FLOATS);
    float* P = wbase + SDF_SMEM_PER_WARP;
    float* sgd = P + BWD_P_FLOATS;
    for (int e = lane; e < BWD_PER_WARP; e += 32) wbase[e] = 0.f;
    __syncthreads();
    for (int e = threadIdx.x; e < SDF_W0_FLOATS; e += blockDim.x) {
        const int j = e / NGLOD_KPAD, k = e - j * NGLOD_KPAD;
        smem[BWD_W2_OFF + ((j >> 1) * NGLOD_KPAD + k) * 2 + (j & 1)] = smem[e];
    }
    __syncthreads();
    const float4* w2v = reinterpret_cast<const float4*>(smem + BWD_W2_OFF);   // [j/2][k/2] -> {W[j][k], W[j+1][k], W[j][k+1], W[j+1][k+1]}

    const float4* w4 = reinterpret_cast<const float4*>(smem);
    const float* sw1 = smem + SDF_SMEM_W1_OFF;
    const float sb1 = smem[SDF_SMEM_B1_OFF];

    // phase-C accumulators: hidden units j = lane + 32*m
    // (packed pairs over k: every FMA below is an FFMA2, each half rounded like the scalar fmaf it replaces)
    uint64_t accW0[4][NGLOD_KPAD / 2];
    float accW1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int k = 0; k < NGLOD_KPAD / 2; ++k) accW0[m][k] = 0ull;
    float acc_b1 = 0.f, acc_loss = 0.f;
    float w1m[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) w1m[m] = sw1[lane + 32 * m];

    const long long gwarp = (long long)blockIdx.x * SDF_WARPS + warp;
    const long long nwarps = (long long)gridDim.x * SDF_WARPS;
    for (long long base = gwarp * 32; base < n; base += nwarps * 32) {
        const long long i = base + lane;
        const bool active = i < n;
        float px = 0.f, py = 0.f, pz = 0.f;
        if (active) { px = __ldg(x + 3 * i); py = __ldg(x + 3 * i + 1); pz = __ldg(x + 3 * i + 2); }
        // ---- A: gather
        warp_gather_tile(net, px, py, pz, active, tile, idx, lane);
        if (!active) {   // keep inactive rows finite and inert
#pragma unroll
            for (int k4 = 0; k4 < NGLOD_KPAD / 4; ++k4)
                *reinterpret_cast<float4*>(tile + lane * NGLOD_KPAD + 4 * k4) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // ---- B: thread-per-query
        float in[NGLOD_KPAD];
#pragma unroll
        for (int k4 = 0; k4 < NGLOD_KPAD / 4; ++k4) {
            const float4 v = *reinterpret_cast<const float4*>(tile + lane * NGLOD_KPAD + 4 * k4);
            in[4 * k4] = v.x; in[4 * k4 + 1] = v.y; in[4 * k4 + 2] = v.z; in[4 * k4 + 3] = v.w;
        }
        float gd = 0.f;
        // pre-activations a_j = W0[j].in for two hidden units per FFMA2 (interleaved copy of W0), same k order as before
        auto pre_pair = [&](int jp, float& a0, float& a1) {
            uint64_t a = 0ull;
#pragma unroll
            for (int k2 = 0; k2 < NGLOD_KPAD / 2; ++k2) {
                const float4 w = w2v[jp * (NGLOD_KPAD / 2) + k2];
                a = f2_fma(f2_pack(w.x, w.y), f2_pack(in[2 * k2], in[2 * k2]), a);
                a = f2_fma(f2_pack(w.z, w.w), f2_pack(in[2 * k2 + 1], in[2 * k2 + 1]), a);
            }
            f2_unpack(a, a0, a1);
        };
        if (FUSED_LOSS) {
            float d = sb1;
#pragma unroll 2
            for (int jp = 0; jp < NGLOD_H / 2; ++jp) {
                float a0, a1;
                pre_pair(jp, a0, a1);
                P[lane * BWD_P_STRIDE + 2 * jp] = a0;
                P[lane * BWD_P_STRIDE + 2 * jp + 1] = a1;
                d = fmaf(sw1[2 * jp], fmaxf(a0, 0.f), d);
                d = fmaf(sw1[2 * jp + 1], fmaxf(a1, 0.f), d);
            }
            if (active) {
                const float diff = d - __ldg(gt + i);
                acc_loss = fmaf(diff * diff, loss_scale, acc_loss);
                gd = 2.f * diff * loss_scale;
            }
        } else {
            if (active) gd = __ldg(grad_out + i);
        }
        // g_in[k] = sum_j W0[j][k] * g_h[j], pairs over k
        uint64_t gin2[NGLOD_F / 2];
        float gx0 = 0.f, gx1 = 0.f, gx2 = 0.f;
#pragma unroll
        for (int k = 0; k < NGLOD_F / 2; ++k) gin2[k] = 0ull;
#pragma unroll 2
        for (int jp = 0; jp < NGLOD_H / 2; ++jp) {
            float apair[2];
            if (FUSED_LOSS) {
                apair[0] = P[lane * BWD_P_STRIDE + 2 * jp];
                apair[1] = P[lane * BWD_P_STRIDE + 2 * jp + 1];
            } else {
                pre_pair(jp, apair[0], apair[1]);
                P[lane * BWD_P_STRIDE + 2 * jp] = apair[0];
                P[lane * BWD_P_STRIDE + 2 * jp + 1] = apair[1];
            }
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int j = 2 * jp + jj;
                const float gh = apair[jj] > 0.f ? gd * sw1[j] : 0.f;
                const uint64_t gh2 = f2_pack(gh, gh);
#pragma unroll
                for (int k4 = 0; k4 < NGLOD_F / 4; ++k4) {
                    const float4 wr = w4[j * (NGLOD_KPAD / 4) + k4];
                    gin2[2 * k4] = f2_fma(f2_pack(wr.x, wr.y), gh2, gin2[2 * k4]);
                    gin2[2 * k4 + 1] = f2_fma(f2_pack(wr.z, wr.w), gh2, gin2[2 * k4 + 1]);
                }
                if (WITH_GX) {
                    const float4 wr = w4[j * (NGLOD_KPAD / 4) + 8];
                    gx0 = fmaf(wr.x, gh, gx0); gx1 = fmaf(wr.y, gh, gx1); gx2 = fmaf(wr.z, gh, gx2);
                }
            }
        }
        float gin[NGLOD_KPAD - 1];
#pragma unroll
        for (int k = 0; k < NGLOD_F / 2; ++k) f2_unpack(gin2[k], gin[2 * k], gin[2 * k + 1]);
        gin[32] = gx0; gin[33] = gx1; gin[34] = gx2;
        sgd[lane] = gd;
        acc_b1 += gd;
        __syncwarp();
        // ---- C: head gradients, lane owns 4 hidden units
#pragma unroll 1
        for (int q = 0; q < 32; ++q) {
            const float gdq = sgd[q];
            if (gdq == 0.f) continue;            // warp-uniform (inactive rows, exact-zero upstream grads)
            float gh[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const float p = P[q * BWD_P_STRIDE + lane + 32 * m];
                gh[m] = p > 0.f ? gdq * w1m[m] : 0.f;
                accW1[m] = fmaf(gdq, fmaxf(p, 0.f), accW1[m]);
            }
#pragma unroll
            for (int k4 = 0; k4 < NGLOD_KPAD / 4; ++k4) {
                const float4 v = *reinterpret_cast<const float4*>(tile + q * NGLOD_KPAD + 4 * k4);
                const uint64_t v01 = f2_pack(v.x, v.y), v23 = f2_pack(v.z, v.w);
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    const uint64_t g2 = f2_pack(gh[m], gh[m]);
                    accW0[m][2 * k4] = f2_fma(g2, v01, accW0[m][2 * k4]);
                    accW0[m][2 * k4 + 1] = f2_fma(g2, v23, accW0[m][2 * k4 + 1]);
                }
            }
        }
        __syncwarp();
        // ---- D: scatter g_feat into the grids (and dL/dx)
#pragma unroll
        for (int k4 = 0; k4 < NGLOD_F / 4; ++k4)
            *reinterpret_cast<float4*>(tile + lane * NGLOD_KPAD + 4 * k4) =
                make_float4(gin[4 * k4], gin[4 * k4 + 1], gin[4 * k4 + 2], gin[4 * k4 + 3]);
        __syncwarp();
        {
            const unsigned live = __ballot_sync(0xffffffffu, active);
            const int n_live = __popc(live);
            const int sub = lane >> 3, c = lane & 7;
            float gx_out[3] = {0.f, 0.f, 0.f};    // valid in the query's own lane after the rounds
            for (int r = 0; r * 4 < n_live; ++r) {
                const int slot = r * 4 + sub;
                const bool valid = slot < n_live;
                const int q = idx[valid ? slot : 0];
                const float qx = __shfl_sync(0xffffffffu, px, q);
                const float qy = __shfl_sync(0xffffffffu, py, q);
                const float qz = __shfl_sync(0xffffffffu, pz, q);
                float sx = 0.f, sy = 0.f, sz = 0.f;
                if (valid) {
                    const float4 g = *reinterpret_cast<const float4*>(tile + q * NGLOD_KPAD + 4 * c);
#pragma unroll
                    for (int l = 0; l < NGLOD_MAX_LODS; ++l) {
                        if (l >= net.num_lods) break;
                        const int R = net.res[l], S = R + 1;
                        const BwdAxis ax = bwd_axis(qx, R), ay = bwd_axis(qy, R), az = bwd_axis(qz, R);
                        float* gg = grad.grids[l];
                        float lx = 0.f, ly = 0.f, lz = 0.f;
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int ix = (k & 1) ? ax.i1 : ax.i0;
                            const int iy = (k & 2) ? ay.i1 : ay.i0;
                            const int iz = (k & 4) ? az.i1 : az.i0;
                            const float wx = (k & 1) ? ax.w1 : ax.w0;
                            const float wy = (k & 2) ? ay.w1 : ay.w0;
                            const float wz = (k & 4) ? az.w1 : az.w0;
                            const int off = ((iz * S + iy) * S + ix) * NGLOD_F + 4 * c;
                            const float w = (wx * wy) * wz;
                            if (gg) red_add_v4(gg + off, g.x * w, g.y * w, g.z * w, g.w * w);
                            if (WITH_GX) {
                                const float4 v = ldg_f4(net.grids[l] + off);
                                const float dot = v.x * g.x + v.y * g.y + v.z * g.z + v.w * g.w;
                                lx += ((k & 1) ? dot : -dot) * (wy * wz);
                                ly += ((k & 2) ? dot : -dot) * (wx * wz);
                                lz += ((k & 4) ? dot : -dot) * (wx * wy);
                            }
                        }
                        if (WITH_GX) { sx = fmaf(lx, ax.mult, sx); sy = fmaf(ly, ay.mult, sy); sz = fmaf(lz, az.mult, sz); }
                    }
                }
                if (WITH_GX) {
                    // reduce the 8 channel-group partials of each sub-warp, hand the sum to the query's lane
#pragma unroll
                    for (int o = 4; o >= 1; o >>= 1) {
                        sx += __shfl_xor_sync(0xffffffffu, sx, o);
                        sy += __shfl_xor_sync(0xffffffffu, sy, o);
                        sz += __shfl_xor_sync(0xffffffffu, sz, o);
                    }
                    if (valid && c == 0) {
                        float* dst = tile + q * NGLOD_KPAD + NGLOD_F;   // xyz columns are free now
                        dst[0] = sx; dst[1] = sy; dst[2] = sz;
                    }
                }
            }
            if (WITH_GX) {
                __syncwarp();
                if (active) {
                    const float* src = tile + lane * NGLOD_KPAD + NGLOD_F;
                    gx_out[0] = src[0] + gin[32]; gx_out[1] = src[1] + gin[33]; gx_out[2] = src[2] + gin[34];
                    grad_x[3 * i] = gx_out[0]; grad_x[3 * i + 1] = gx_out[1]; grad_x[3 * i + 2] = gx_out[2];
                }
            }
        }
        __syncwarp();
    }

    // ---- flush head gradients: registers -> CTA accumulator in smem -> one RED per CTA per element
    __syncthreads();
    float* cta_acc = smem + SDF_SMEM_WARP_OFF;       // per-warp regions are dead now
    for (int e = threadIdx.x; e < BWD_ACC_FLOATS; e += blockDim.x) cta_acc[e] = 0.f;
    __syncthreads();
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const int j = lane + 32 * m;
#pragma unroll
        for (int k = 0; k < NGLOD_KPAD / 2; ++k) {
            float a0, a1;
            f2_unpack(accW0[m][k], a0, a1);
            atomicAdd(cta_acc + j * NGLOD_KPAD + 2 * k, a0);
            atomicAdd(cta_acc + j * NGLOD_KPAD + 2 * k + 1, a1);
        }
        atomicAdd(cta_acc + SDF_W0_FLOATS + j, accW1[m]);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        acc_b1 += __shfl_xor_sync(0xffffffffu, acc_b1, o);
        acc_loss += __shfl_xor_sync(0xffffffffu, acc_loss, o);
    }
    if (lane == 0) {
        atomicAdd(cta_acc + SDF_W0_FLOATS + NGLOD_H, acc_b1);
        atomicAdd(cta_acc + SDF_W0_FLOATS + NGLOD_H + 1, acc_loss);
    }
    __syncthreads();
    const int in_dim = net.pos_invariant ? NGLOD_F : NGLOD_F + 3;
    for (int e = threadIdx.x; e < SDF_W0_FLOATS; e += blockDim.x) {
        const int j = e / NGLOD_KPAD, k = e - j * NGLOD_KPAD;
        const float v = cta_acc[e];
        if (k < NGLOD_F) {
            if (grad.w0) atomicAdd(grad.w0 + j * in_dim + (net.pos_invariant ? k : k + 3), v);
        } else if (k < NGLOD_F + 3) {
            if (grad.w0 && !net.pos_invariant) atomicAdd(grad.w0 + j * in_dim + (k - NGLOD_F), v);
        } else {
            if (grad.b0) atomicAdd(grad.b0 + j, v);
        }
    }
    for (int e = threadIdx.x; e < NGLOD_H; e += blockDim.x)
        if (grad.w1) atomicAdd(grad.w1 + e, cta_acc[SDF_W0_FLOATS + e]);
    if (threadIdx.x == 0) {
        if (grad.b1) atomicAdd(grad.b1, cta_acc[SDF_W0_FLOATS + NGLOD_H]);
        if (FUSED_LOSS && loss_out) atomicAdd(loss_out, cta_acc[SDF_W0_FLOATS + NGLOD_H + 1]);
    }
}

// ---- transpose of the prefix sum: push dL/d(summed grid of level l) down to the LOD grids, level by level.
// restrict: Tc[c] += sum over fine nodes n in the support of coarse node c's hat function of  w(n, c) * Tf[n],
//           w = prod_axis (1 - |n_a - k c_a| / k), k = Rf / Rc  -- the weights nglod_build_summed_grid used, transposed.
// KT = the ratio Rf / Rc when it is the octree's 2 (27 taps, fully unrolled: 27 independent loads in flight -- the runtime
// loops walked them one L2 latency at a time, ~10 us even for the 9^3 level), 0 = any ratio (runtime loops).
template <int KT>
__global__ void __launch_bounds__(256)
restrict_add_kernel(const float4* __restrict__ Tf, const int Rf, float4* __restrict__ Tc, const int Rc, const int n_chunks) {
    const int Sc = Rc + 1, Sf = Rf + 1, k = KT ? KT : Rf / Rc;
    const float inv_k = 1.f / (float)k;           // exact for the power-of-two ratios of an octree
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_chunks; e += gridDim.x * blockDim.x) {
        const int c = e & 7;
        int node = e >> 3;
        const int cx = node % Sc; node /= Sc;
        const int cy = node % Sc;
        const int cz = node / Sc;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if constexpr (KT == 2) {
            float4 v[27];
            float w[27];
#pragma unroll
            for (int t = 0; t < 27; ++t) {
                const int dx = t % 3 - 1, dy = (t / 3) % 3 - 1, dz = t / 9 - 1;
                const int fx = 2 * cx + dx, fy = 2 * cy + dy, fz = 2 * cz + dz;
                const bool ok = fx >= 0 && fx <= Rf && fy >= 0 && fy <= Rf && fz >= 0 && fz <= Rf;
                w[t] = ok ? (dx ? 0.5f : 1.f) * (dy ? 0.5f : 1.f) * (dz ? 0.5f : 1.f) : 0.f;
                v[t] = ok ? __ldg(Tf + ((long long)(fz * Sf + fy) * Sf + fx) * 8 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int dz = 0; dz < 3; ++dz)           // same order as the runtime loops: z outer, x inner
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        const int t = dz * 9 + dy * 3 + dx;
                        if (w[t] != 0.f) {
                            acc.x = fmaf(v[t].x, w[t], acc.x); acc.y = fmaf(v[t].y, w[t], acc.y);
                            acc.z = fmaf(v[t].z, w[t], acc.z); acc.w = fmaf(v[t].w, w[t], acc.w);
                        }
                    }
        } else {
            for (int dz = -(k - 1); dz <= k - 1; ++dz) {
                const int fz = cz * k + dz;
                if (fz < 0 || fz > Rf) continue;
                const float wz = 1.f - (float)abs(dz) * inv_k;
                for (int dy = -(k - 1); dy <= k - 1; ++dy) {
                    const int fy = cy * k + dy;
                    if (fy < 0 || fy > Rf) continue;
                    const float wzy = wz * (1.f - (float)abs(dy) * inv_k);
                    for (int dx = -(k - 1); dx <= k - 1; ++dx) {
                        const int fx = cx * k + dx;
                        if (fx < 0 || fx > Rf) continue;
                        const float w = wzy * (1.f - (float)abs(dx) * inv_k);
                        const float4 v = __ldg(Tf + ((long long)(fz * Sf + fy) * Sf + fx) * 8 + c);
                        acc.x = fmaf(v.x, w, acc.x); acc.y = fmaf(v.y, w, acc.y); acc.z = fmaf(v.z, w, acc.z); acc.w = fmaf(v.w, w, acc.w);
                    }
                }
            }
        }
        float4 t = Tc[e];
        t.x += acc.x; t.y += acc.y; t.z += acc.z; t.w += acc.w;
        Tc[e] = t;
    }
}

// g += T (if g), T = 0
__global__ void __launch_bounds__(256)
accumulate_and_clear_kernel(float4* __restrict__ T, float4* __restrict__ g, const long long n4) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (long long)gridDim.x * blockDim.x) {
        const float4 t = T[e];
        if (g) {
            float4 v = g[e];
            v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
            g[e] = v;
        }
        T[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// target += sum of the private scatter copies; the copies are zero again afterwards
// 8 lanes per float4 element, each over every 8th copy, then a shuffle reduction (one thread per element walked up to 64
// copies serially: 21 us for 1000 elements)
__global__ void __launch_bounds__(256)
fold_copies_kernel(float4* __restrict__ priv, const int copies, const long long n4, float4* __restrict__ target) {
    const long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int part = threadIdx.x & 7;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e < n4)
        for (int c = part; c < copies; c += 8) {
            const float4 v = priv[c * n4 + e];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            priv[c * n4 + e] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
    for (int o = 4; o >= 1; o >>= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
        acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
    }
    if (e < n4 && part == 0 && target) {
        float4 t = target[e];
        t.x += acc.x; t.y += acc.y; t.z += acc.z; t.w += acc.w;
        target[e] = t;
    }
}

int restrict_cascade(const nglod_net_t* net, int lod, const nglod_net_grad_t* grad, cudaStream_t st) {
    const long long cap = (long long)nglod_sm_count() * 16;
    for (int l = lod; l >= 0; --l) {
        const long long S = net->grid_res[l] + 1;
        const long long n4 = S * S * S * 8;
        if (l > 0) {
            const long long Sc = net->grid_res[l - 1] + 1;
            const long long nc = Sc * Sc * Sc * 8;
            long long blocks = (nc + 255) / 256;
            if (blocks > cap) blocks = cap;
            if (net->grid_res[l] == 2 * net->grid_res[l - 1])
                restrict_add_kernel<2><<<(int)blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(grad->summed[l]), net->grid_res[l],
                                                                    reinterpret_cast<float4*>(grad->summed[l - 1]), net->grid_res[l - 1], (int)nc);
            else
                restrict_add_kernel<0><<<(int)blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(grad->summed[l]), net->grid_res[l],
                                                                    reinterpret_cast<float4*>(grad->summed[l - 1]), net->grid_res[l - 1], (int)nc);
        }
        if (grad->summed[l] == grad->grids[l]) continue;      // aliased: the level's gradient was accumulated in place
        long long blocks = (n4 + 255) / 256;
        if (blocks > cap) blocks = cap;
        accumulate_and_clear_kernel<<<(int)blocks, 256, 0, st>>>(reinterpret_cast<float4*>(grad->summed[l]),
                                                                 reinterpret_cast<float4*>(grad->grids[l]), n4);
    }
    return (int)cudaGetLastError();
}

// the single-grid path needs the summed grid of this LOD, zeroed scratch for every level down the chain, nesting grids
bool use_summed_backward(const nglod_net_t* net, int lod, const nglod_net_grad_t* grad) {
    if (!net->summed[lod] || !grad) return false;
    for (int i = 0; i <= lod; ++i) {
        if (!grad->summed[i] || (reinterpret_cast<uintptr_t>(grad->summed[i]) & 15u)) return false;
        if (i > 0 && net->grid_res[i] % net->grid_res[i - 1] != 0) return false;
    }
    return true;
}

template <bool FUSED_LOSS, bool WITH_GX>
int launch_backward(const nglod_net_t* net, int lod, const nglod_net_grad_t* grad, const float* x, int64_t n,
                    const float* grad_out, const float* gt, float loss_scale, float* grad_x, float* loss_out,
                    cudaStream_t st, bool cascade = true) {
    const bool single = use_summed_backward(net, lod, grad);
    const NetDev nd = single ? nglod_make_netdev_infer(net, lod, /*allow_half=*/false) : nglod_make_netdev(net, lod);
    GradDev gdv;
    for (int i = 0; i < NGLOD_MAX_LODS; ++i) gdv.grids[i] = (grad && i <= lod && !single) ? grad->grids[i] : nullptr;
    if (single) gdv.grids[0] = grad->summed[lod];
    gdv.w0 = grad ? grad->w0[lod] : nullptr;
    gdv.b0 = grad ? grad->b0[lod] : nullptr;
    gdv.w1 = grad ? grad->w1[lod] : nullptr;
    gdv.b1 = grad ? grad->b1[lod] : nullptr;
    if constexpr (!WITH_GX) {
        // tcgen05 kernel (sdf_backward_tc.cu); the FP32 kernel above serves dL/dx
        if (single) {
            // a grid of <= 729 nodes: scatter into private copies (enough of them to spread the batch over ~5000 nodes)
            const long long grid_floats = (long long)(nd.res[0] + 1) * (nd.res[0] + 1) * (nd.res[0] + 1) * NGLOD_F;
            if (grad->scatter_scratch && !(reinterpret_cast<uintptr_t>(grad->scatter_scratch) & 15u) && nd.res[0] <= 8 &&
                n >= 4 * grid_floats) {
                long long copies = (160000 + grid_floats - 1) / grid_floats;
                if (copies > grad->scatter_scratch_floats / grid_floats) copies = grad->scatter_scratch_floats / grid_floats;
                if (copies > 64) copies = 64;
                if (copies > 1) { gdv.priv = grad->scatter_scratch; gdv.priv_copies = (int)copies; gdv.priv_stride = (int)grid_floats; }
            }
        }
        if (int e = nglod_launch_sdf_backward_tc(nd, gdv, x, (long long)n, grad_out, gt, loss_scale, loss_out, FUSED_LOSS, st)) return e;
        if (gdv.priv) {
            const long long n4 = (long long)gdv.priv_stride / 4;
            fold_copies_kernel<<<(int)((n4 * 8 + 255) / 256), 256, 0, st>>>(reinterpret_cast<float4*>(gdv.priv), gdv.priv_copies, n4,
                                                                        reinterpret_cast<float4*>(gdv.grids[0]));
            if (int e = (int)cudaGetLastError()) return e;
        }
        return (single && cascade) ? restrict_cascade(net, lod, grad, st) : 0;
    }
    long long grid = nglod_sm_count();
    auto kern = sdf_backward_kernel<FUSED_LOSS, WITH_GX>;
    NGLOD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM_BYTES));
    const long long want = (n + SDF_THREADS - 1) / SDF_THREADS;
    if (want < grid) grid = want;
    kern<<<(int)grid, SDF_THREADS, BWD_SMEM_BYTES, st>>>(nd, gdv, x, (long long)n, grad_out, gt, loss_scale, grad_x,
                                                         loss_out);
    if (int e = (int)cudaGetLastError()) return e;
    return (single && cascade) ? restrict_cascade(net, lod, grad, st) : 0;
}

}  // namespace

extern "C" int nglod_sdf_backward(const nglod_net_t* net, int32_t lod, const float* x, int64_t n,
                                  const float* grad_out, const nglod_net_grad_t* grad, float* grad_x,
                                  void* stream) {
    if (int e = nglod_check_net(net, lod)) return e;
    if (n < 0 || (n > 0 && (!x || !grad_out))) return NGLOD_EINVAL;
    if (n == 0) return 0;
    for (int i = 0; grad && i <= lod; ++i)
        if (grad->grids[i] && (reinterpret_cast<uintptr_t>(grad->grids[i]) & 15u)) return NGLOD_EINVAL;
    if (grad_x)
        return launch_backward<false, true>(net, lod, grad, x, n, grad_out, nullptr, 0.f, grad_x, nullptr,
                                            (cudaStream_t)stream);
    return launch_backward<false, false>(net, lod, grad, x, n, grad_out, nullptr, 0.f, nullptr, nullptr,
                                         (cudaStream_t)stream);
}

template <bool FUSED_LOSS>
static int launch_sparse_backward(const nglod_sparse_net_t* net, int32_t lod, const float* x, const int32_t* pidx, int64_t n,
                                  const float* grad_out, const float* gt, float loss_scale, float* grad_corner_feats,
                                  float* gw0, float* gb0, float* gw1, float* gb1, float* loss_out, void* stream) {
    SparseBwd sp;
    if (int e = make_sparse_dev(net, lod, sp.sn, /*allow_summed=*/false)) return e;
    if (n < 0 || (n > 0 && (!x || !pidx || !(FUSED_LOSS ? gt : grad_out)))) return NGLOD_EINVAL;
    if (reinterpret_cast<uintptr_t>(grad_corner_feats) & 15u) return NGLOD_EINVAL;
    if (n == 0) return 0;
    sp.pidx = pidx;
    sp.grad_cf = grad_corner_feats;
    GradDev gdv;
    for (int i = 0; i < NGLOD_MAX_LODS; ++i) gdv.grids[i] = nullptr;
    gdv.w0 = gw0; gdv.b0 = gb0; gdv.w1 = gw1; gdv.b1 = gb1;
    return nglod_launch_sdf_backward_tc(sp.sn.dec, gdv, x, (long long)n, grad_out, gt, loss_scale, loss_out, FUSED_LOSS,
                                        (cudaStream_t)stream, &sp);
}

extern "C" int nglod_sparse_sdf_backward(const nglod_sparse_net_t* net, int32_t lod, const float* x, const int32_t* pidx,
                                         int64_t n, const float* grad_out, float* grad_corner_feats, float* gw0,
                                         float* gb0, float* gw1, float* gb1, void* stream) {
    return launch_sparse_backward<false>(net, lod, x, pidx, n, grad_out, nullptr, 0.f, grad_corner_feats, gw0, gb0, gw1, gb1,
                                         nullptr, stream);
}

extern "C" int nglod_sparse_sdf_train_step(const nglod_sparse_net_t* net, int32_t lod, const float* x, const int32_t* pidx,
                                           const float* gt, int64_t n, float loss_scale, float* grad_corner_feats,
                                           float* gw0, float* gb0, float* gw1, float* gb1, float* loss_out, void* stream) {
    return launch_sparse_backward<true>(net, lod, x, pidx, n, nullptr, gt, loss_scale, grad_corner_feats, gw0, gb0, gw1, gb1,
                                        loss_out, stream);
}

extern "C" int nglod_sdf_train_step(const nglod_net_t* net, uint32_t lod_mask, const float* x, const float* gt,
                                    int64_t n, float loss_scale, const nglod_net_grad_t* grad, float* loss_out,
                                    void* stream) {
    if (!net || !grad) return NGLOD_EINVAL;
    if (n < 0 || (n > 0 && (!x || !gt))) return NGLOD_EINVAL;
    if (n == 0) return 0;
    // every head scatters into the scratch of its own level; ONE cascade from the top head then pushes the lot down
    int top_single = -1;
    for (int l = 0; l < net->num_lods; ++l) {
        if (!(lod_mask & (1u << l))) continue;
        if (int e = nglod_check_net(net, l)) return e;
        for (int i = 0; i <= l; ++i)
            if (grad->grids[i] && (reinterpret_cast<uintptr_t>(grad->grids[i]) & 15u)) return NGLOD_EINVAL;
        if (use_summed_backward(net, l, grad)) top_single = l;
        float* lo = (loss_out && (lod_mask & NGLOD_LOSS_PER_LOD)) ? loss_out + l : loss_out;
        if (int e = launch_backward<true, false>(net, l, grad, x, n, nullptr, gt, loss_scale, nullptr, lo,
                                                 (cudaStream_t)stream, /*cascade=*/false))
            return e;
    }
    if (top_single >= 0) return restrict_cascade(net, top_single, grad, (cudaStream_t)stream);
    return 0;
}
