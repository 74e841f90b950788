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

// ------------------------------------------------------------------------------------------------------------------
// Backward, second generation (no dL/dx): 16 warps per CTA, head gradients on the tensor cores.
//
// The first-generation kernel above keeps dW0|db0 as 144 register accumulators per lane (255 registers -> 8 warps per
// SM, latency-bound at 39 % issue).  Here the only state a query leaves behind for the head gradients is its upstream
// gradient g_d and the 128 ReLU mask bits:
//     T[h][k]  = sum_q g_d(q) [pre_qh > 0] in_qk           (k = 0..35, in_q35 = 1)       <- one GEMM over the queries
//     dW0[h][k] = w1[h] T[h][k],  db0[h] = w1[h] T[h][35],  dW1[h] = sum_k W0ext[h][k] T[h][k]   (pre = W0ext . in)
// and T is accumulated CTA-wide with mma.sync.m16n8k8 TF32 (rna-rounded operands, fp32 accumulate): warp w owns hidden
// units [16 (w%8), +16) and the queries of warps [8 (w/8), +8) of the batch -> 20 accumulator registers.  The two
// per-query GEMMs run on mma.sync as well, 16 queries at a time per warp: the forward recompute pre = in . W0ext^T in
// 3xTF32 (d feeds the loss gradient, so it keeps fp32-level accuracy), dL/d(features) = g_h . W0 likewise (its terms cancel heavily),
// with its A fragments built in registers from the mask bits of the pre fragments.  128 registers, 16 warps per SM.
#define BW2_WARPS 16
#define BW2_THREADS (BW2_WARPS * 32)
#define BW2_PER_WARP (SDF_SMEM_PER_WARP + 32 * 4 + 32)           // tile, idx, mask words [32][4], gd[32]
#define BW2_WLO_OFF (SDF_SMEM_WARP_OFF + BW2_WARPS * BW2_PER_WARP)   // TF32 low parts of W0ext (the staged copy keeps the high parts)
#define BW2_SMEM_BYTES ((BW2_WLO_OFF + SDF_W0_FLOATS) * 4)
#ifndef BW2_SMEM_GRID_MAX           // -DBW2_SMEM_GRID_MAX=23328 also takes R = 8 (93 KB): untested, see profiles/NEXT.md
#define BW2_SMEM_GRID_MAX 4000      // floats: a 5^3 x 32 grid (R = 4) accumulated per CTA in shared memory
#endif

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r;
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// SPARSE: the features come from (and their gradients go to) the corner rows of a sparse octree model -- gather and
// scatter walk the query's parent chain (sparse_core.cuh); `net` is then sp.sn.dec, the decoder of the LOD.
struct SparseBwd {
    SparseDev sn;
    const int* pidx;        // [n] voxel index within the LOD's level
    float* grad_cf;         // [NC, F]
};

template <bool FUSED_LOSS, bool SPARSE>
__global__ void __launch_bounds__(BW2_THREADS, 1)
sdf_backward_mma_kernel(const NetDev net, const GradDev grad, const float* __restrict__ x, const long long n,
                        const float* __restrict__ grad_out, const float* __restrict__ gt, const float loss_scale,
                        float* __restrict__ loss_out, const SparseBwd sp, const int lpw, const int smem_grid_floats) {
    // lpw = queries per warp per batch: 32, or 8 when the whole call is too small to give every SM a 512-query batch
    // (a batch is then 128 queries: the kernel's serial phases are 3-4x shorter and 4x as many CTAs share the work)
    extern __shared__ __align__(16) float smem[];
    sdf_stage_weights(net, smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* wbase = smem + SDF_SMEM_WARP_OFF + warp * BW2_PER_WARP;
    float* tile = wbase;
    int* idx = reinterpret_cast<int*>(wbase + SDF_TILE_FLOATS);
    uint32_t* maskw = reinterpret_cast<uint32_t*>(wbase + SDF_SMEM_PER_WARP);
    float* sgd = wbase + SDF_SMEM_PER_WARP + 32 * 4;
    for (int e = lane; e < BW2_PER_WARP; e += 32) wbase[e] = 0.f;
    // A grid of R = 4 has 125 nodes: half a million queries scattering into its 16 KB serialise in the L2 atomic units (the
    // LOD-0 launch of a training step took 0.79 ms against 0.37 ms for every other level).  Such a grid (smem_grid_floats
    // > 0, level 0 only) is accumulated in shared memory and flushed once per CTA.
    float* sgrid = smem + BW2_SMEM_BYTES / 4;
    for (int e = threadIdx.x; e < smem_grid_floats; e += blockDim.x) sgrid[e] = 0.f;
    __syncthreads();
    // split the staged W0ext once: smem[e] = TF32 high part, wlo[e] = TF32 of the remainder (B fragments of the 3xTF32 GEMMs)
    float* wlo = smem + BW2_WLO_OFF;
    for (int e = threadIdx.x; e < SDF_W0_FLOATS; e += blockDim.x) {
        const float w = smem[e];
        const float hi = __uint_as_float(to_tf32(w));
        smem[e] = hi;
        wlo[e] = __uint_as_float(to_tf32(w - hi));
    }
    __syncthreads();
    const float* sw1 = smem + SDF_SMEM_W1_OFF;
    const float sb1 = smem[SDF_SMEM_B1_OFF];

    // phase-C ownership: hidden units [16 mt, 16 mt + 16), queries of warps [8 qh, 8 qh + 8) of every CTA batch
    const int mt = warp & 7, qh = warp >> 3;
    const int g = lane >> 2, t = lane & 3;
    float accT[5][4];
#pragma unroll
    for (int nt = 0; nt < 5; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) accT[nt][i] = 0.f;
    float acc_b1 = 0.f, acc_loss = 0.f;

    const long long batch = (long long)BW2_WARPS * lpw;
    for (long long base0 = (long long)blockIdx.x * batch; base0 < n; base0 += (long long)gridDim.x * batch) {
        const long long i = base0 + warp * lpw + lane;
        bool active = lane < lpw && i < n;
        int pv = 0;
        if constexpr (SPARSE) {      // pidx < 0 (point outside the octree): the row is inert -- no gather, no loss, no gradient
            if (active) pv = __ldg(sp.pidx + i);
            active = active && pv >= 0;
        }
        const unsigned live_rows = __ballot_sync(0xffffffffu, active);
        float px = 0.f, py = 0.f, pz = 0.f;
        if (active) { px = __ldg(x + 3 * i); py = __ldg(x + 3 * i + 1); pz = __ldg(x + 3 * i + 2); }
        // ---- A: gather
        int vrow = 0;
        if constexpr (SPARSE) {
            vrow = sp.sn.vox_off + (active ? pv : 0);
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
                if (valid) *reinterpret_cast<float4*>(tile + q * NGLOD_KPAD + 4 * c) = sparse_gather4(sp.sn, qx, qy, qz, qv, c);
            }
            __syncwarp();
        } else {
            warp_gather_tile(net, px, py, pz, active, tile, idx, lane);
        }
        if (!active) {   // keep inactive rows finite and inert
#pragma unroll
            for (int k4 = 0; k4 < NGLOD_KPAD / 4; ++k4)
                *reinterpret_cast<float4*>(tile + lane * NGLOD_KPAD + 4 * k4) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // ---- B: the decoder's two GEMMs on the tensor cores (mma.sync m16n8k8 TF32), 16 queries (one m-tile) at a time:
        //      pre[q][h] = in[q] . W0ext[h]   3xTF32 (A_lo B_hi + A_hi B_lo + A_hi B_hi: d feeds the loss gradient)
        //      g_in[q][f] = sum_h g_h[q][h] W0[h][f]   single TF32 pass; its A fragments are built in registers from the
        //      ReLU mask bits of the pre fragments (hidden units of a k-step permuted identically in A and B)
        __syncwarp();
        float ginf[2][4][4];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            if (16 * m >= lpw) {                                      // no query rows in this m-tile (warp-uniform)
#pragma unroll
                for (int nf = 0; nf < 4; ++nf)
#pragma unroll
                    for (int e = 0; e < 4; ++e) ginf[m][nf][e] = 0.f;
                continue;
            }
            const int rA = 16 * m + g, rB = rA + 8;                   // this lane's two query rows of the m-tile
            uint32_t ahi[5][4], alo[5][4];
#pragma unroll
            for (int ks = 0; ks < 5; ++ks) {
                const int k0 = 8 * ks + t, k1 = k0 + 4;
                const float v[4] = {tile[rA * NGLOD_KPAD + k0], tile[rB * NGLOD_KPAD + k0],
                                    k1 < NGLOD_KPAD ? tile[rA * NGLOD_KPAD + k1] : 0.f,
                                    k1 < NGLOD_KPAD ? tile[rB * NGLOD_KPAD + k1] : 0.f};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    ahi[ks][e] = to_tf32(v[e]);
                    alo[ks][e] = to_tf32(v[e] - __uint_as_float(ahi[ks][e]));
                }
            }
            uint32_t bitsA = 0u, bitsB = 0u;                          // bit 2n+e: pre(row, hidden 8n + 2t + e) > 0
            float dA = 0.f, dB = 0.f;
#pragma unroll
            for (int nc = 0; nc < 2; ++nc) {
                float acc[8][4];
#pragma unroll
                for (int n8 = 0; n8 < 8; ++n8)
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[n8][e] = 0.f;
#pragma unroll
                for (int ks = 0; ks < 5; ++ks) {
                    const int k0 = 8 * ks + t, k1 = k0 + 4;
#pragma unroll
                    for (int n8 = 0; n8 < 8; ++n8) {
                        const int h = 8 * (8 * nc + n8) + g;
                        uint32_t bhi[2], blo[2];
                        bhi[0] = __float_as_uint(smem[h * NGLOD_KPAD + k0]);
                        blo[0] = __float_as_uint(wlo[h * NGLOD_KPAD + k0]);
                        bhi[1] = k1 < NGLOD_KPAD ? __float_as_uint(smem[h * NGLOD_KPAD + k1]) : 0u;
                        blo[1] = k1 < NGLOD_KPAD ? __float_as_uint(wlo[h * NGLOD_KPAD + k1]) : 0u;
                        mma_tf32_16x8x8(acc[n8], alo[ks], bhi);
                        mma_tf32_16x8x8(acc[n8], ahi[ks], blo);
                        mma_tf32_16x8x8(acc[n8], ahi[ks], bhi);
                    }
                }
#pragma unroll
                for (int n8 = 0; n8 < 8; ++n8) {
                    const int n = 8 * nc + n8;
                    const float2 w1p = *reinterpret_cast<const float2*>(sw1 + 8 * n + 2 * t);
                    if (FUSED_LOSS) {
                        dA = fmaf(w1p.x, fmaxf(acc[n8][0], 0.f), dA); dA = fmaf(w1p.y, fmaxf(acc[n8][1], 0.f), dA);
                        dB = fmaf(w1p.x, fmaxf(acc[n8][2], 0.f), dB); dB = fmaf(w1p.y, fmaxf(acc[n8][3], 0.f), dB);
                    }
                    bitsA |= ((acc[n8][0] > 0.f ? 1u : 0u) | (acc[n8][1] > 0.f ? 2u : 0u)) << (2 * n);
                    bitsB |= ((acc[n8][2] > 0.f ? 1u : 0u) | (acc[n8][3] > 0.f ? 2u : 0u)) << (2 * n);
                }
            }
            // d of the two rows (sum over the 4 lanes that share a row), upstream gradients
            const long long iA = base0 + warp * lpw + rA, iB = base0 + warp * lpw + rB;
            // a row takes part iff its lane is active (in range and, on the sparse path, inside the octree): inert rows add
            // no loss, no upstream gradient, hence no parameter gradient
            const bool okA = (live_rows >> rA) & 1u, okB = (live_rows >> rB) & 1u;
            float gdA = 0.f, gdB = 0.f;
            if (FUSED_LOSS) {
                dA += __shfl_xor_sync(0xffffffffu, dA, 1); dA += __shfl_xor_sync(0xffffffffu, dA, 2);
                dB += __shfl_xor_sync(0xffffffffu, dB, 1); dB += __shfl_xor_sync(0xffffffffu, dB, 2);
                if (okA) { const float diff = (dA + sb1) - __ldg(gt + iA); gdA = 2.f * diff * loss_scale;
                           if (t == 0) acc_loss = fmaf(diff * diff, loss_scale, acc_loss); }
                if (okB) { const float diff = (dB + sb1) - __ldg(gt + iB); gdB = 2.f * diff * loss_scale;
                           if (t == 0) acc_loss = fmaf(diff * diff, loss_scale, acc_loss); }
            } else {
                if (okA) gdA = __ldg(grad_out + iA);
                if (okB) gdB = __ldg(grad_out + iB);
            }
            if (t == 0) { acc_b1 += gdA + gdB; sgd[rA] = gdA; sgd[rB] = gdB; }
            // ReLU mask words for phase C: word w = hidden [32w, 32w+32); this lane holds the bit pairs (2t, 2t+1) of
            // every 8-wide block; OR over the 4 lanes of the row
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                uint32_t wa = 0u, wb = 0u;
#pragma unroll
                for (int nn = 0; nn < 4; ++nn) {
                    wa |= ((bitsA >> (8 * w + 2 * nn)) & 3u) << (8 * nn + 2 * t);
                    wb |= ((bitsB >> (8 * w + 2 * nn)) & 3u) << (8 * nn + 2 * t);
                }
                wa |= __shfl_xor_sync(0xffffffffu, wa, 1); wa |= __shfl_xor_sync(0xffffffffu, wa, 2);
                wb |= __shfl_xor_sync(0xffffffffu, wb, 1); wb |= __shfl_xor_sync(0xffffffffu, wb, 2);
                if (t == 0) { maskw[rA * 4 + w] = wa; maskw[rB * 4 + w] = wb; }
            }
            // g_in of the m-tile
#pragma unroll
            for (int nf = 0; nf < 4; ++nf)
#pragma unroll
                for (int e = 0; e < 4; ++e) ginf[m][nf][e] = 0.f;
#pragma unroll 4
            for (int n = 0; n < 16; ++n) {
                const int h0 = 8 * n + 2 * t;                                  // A/B "column t" = h0, "column t+4" = h0 + 1
                const float2 w1p = *reinterpret_cast<const float2*>(sw1 + h0);
                const float av[4] = {((bitsA >> (2 * n)) & 1u) ? gdA * w1p.x : 0.f, ((bitsB >> (2 * n)) & 1u) ? gdB * w1p.x : 0.f,
                                     ((bitsA >> (2 * n + 1)) & 1u) ? gdA * w1p.y : 0.f, ((bitsB >> (2 * n + 1)) & 1u) ? gdB * w1p.y : 0.f};
                uint32_t a[4], al[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) { a[e] = to_tf32(av[e]); al[e] = to_tf32(av[e] - __uint_as_float(a[e])); }
#pragma unroll
                for (int nf = 0; nf < 4; ++nf) {
                    uint32_t bb[2], bl[2];
                    bb[0] = __float_as_uint(smem[h0 * NGLOD_KPAD + 8 * nf + g]);
                    bl[0] = __float_as_uint(wlo[h0 * NGLOD_KPAD + 8 * nf + g]);
                    bb[1] = __float_as_uint(smem[(h0 + 1) * NGLOD_KPAD + 8 * nf + g]);
                    bl[1] = __float_as_uint(wlo[(h0 + 1) * NGLOD_KPAD + 8 * nf + g]);
                    mma_tf32_16x8x8(ginf[m][nf], al, bb);        // 3xTF32: the terms of g_in cancel heavily, a single pass
                    mma_tf32_16x8x8(ginf[m][nf], a, bl);         // left 2.4e-4 of max|grad| on the grid gradients
                    mma_tf32_16x8x8(ginf[m][nf], a, bb);
                }
            }
        }
        __syncthreads();
        // ---- C: T[h][k] += sum_q g_d(q) mask(q,h) in(q,k) on the tensor cores
#pragma unroll 2
        for (int ks = 0; ks < 32; ++ks) {
            const int wq = qh * 8 + (ks >> 2);                        // warp that owns these 8 queries
            const int r0 = (ks & 3) * 8;                              // their first row in that warp's tile
            const float* wb = smem + SDF_SMEM_WARP_OFF + wq * BW2_PER_WARP;
            const uint32_t* mq = reinterpret_cast<const uint32_t*>(wb + SDF_SMEM_PER_WARP);
            const float* gq = wb + SDF_SMEM_PER_WARP + 32 * 4;
            const float gd0 = gq[r0 + t], gd1 = gq[r0 + t + 4];
            const uint32_t m0 = mq[(r0 + t) * 4 + (mt >> 1)] >> ((mt & 1) * 16);
            const uint32_t m1 = mq[(r0 + t + 4) * 4 + (mt >> 1)] >> ((mt & 1) * 16);
            uint32_t a[4];
            a[0] = to_tf32(((m0 >> g) & 1u) ? gd0 : 0.f);
            a[1] = to_tf32(((m0 >> (g + 8)) & 1u) ? gd0 : 0.f);
            a[2] = to_tf32(((m1 >> g) & 1u) ? gd1 : 0.f);
            a[3] = to_tf32(((m1 >> (g + 8)) & 1u) ? gd1 : 0.f);
            if (!__any_sync(0xffffffffu, (gd0 != 0.f) | (gd1 != 0.f))) continue;       // 8 inert queries (warp-uniform)
#pragma unroll
            for (int nt = 0; nt < 5; ++nt) {
                uint32_t b[2];
                const int k = 8 * nt + g;
                const bool ok = k < NGLOD_KPAD;
                b[0] = to_tf32(ok ? wb[(r0 + t) * NGLOD_KPAD + k] : 0.f);
                b[1] = to_tf32(ok ? wb[(r0 + t + 4) * NGLOD_KPAD + k] : 0.f);
                mma_tf32_16x8x8(accT[nt], a, b);
            }
        }
        __syncthreads();
        // ---- D: scatter g_feat into the grids (fragments -> the warp's tile rows -> 8 lanes per corner line)
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int nf = 0; nf < 4; ++nf) {
                *reinterpret_cast<float2*>(tile + (16 * m + g) * NGLOD_KPAD + 8 * nf + 2 * t) = make_float2(ginf[m][nf][0], ginf[m][nf][1]);
                *reinterpret_cast<float2*>(tile + (16 * m + g + 8) * NGLOD_KPAD + 8 * nf + 2 * t) = make_float2(ginf[m][nf][2], ginf[m][nf][3]);
            }
        __syncwarp();
        {
            const unsigned live = __ballot_sync(0xffffffffu, active);
            const int n_live = __popc(live);
            const int sub = lane >> 3, c = lane & 7;
            for (int r = 0; r * 4 < n_live; ++r) {
                const int slot = r * 4 + sub;
                const bool valid = slot < n_live;
                const int q = idx[valid ? slot : 0];
                const float qx = __shfl_sync(0xffffffffu, px, q);
                const float qy = __shfl_sync(0xffffffffu, py, q);
                const float qz = __shfl_sync(0xffffffffu, pz, q);
                const int qv = __shfl_sync(0xffffffffu, vrow, q);
                if (valid && SPARSE) {
                    const float4 gq4 = *reinterpret_cast<const float4*>(tile + q * NGLOD_KPAD + 4 * c);
                    if (sp.grad_cf) sparse_scatter4(sp.sn, sp.grad_cf, qx, qy, qz, qv, c, gq4);
                }
                if (valid && !SPARSE) {
                    const float4 gq4 = *reinterpret_cast<const float4*>(tile + q * NGLOD_KPAD + 4 * c);
#pragma unroll
                    for (int l = 0; l < NGLOD_MAX_LODS; ++l) {
                        if (l >= net.num_lods) break;
                        const int R = net.res[l], S = R + 1;
                        const BwdAxis ax = bwd_axis(qx, R), ay = bwd_axis(qy, R), az = bwd_axis(qz, R);
                        float* gg = grad.grids[l];
                        if (!gg) continue;
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
                            if (l == 0 && smem_grid_floats > 0) {
                                atomicAdd(sgrid + off, gq4.x * w); atomicAdd(sgrid + off + 1, gq4.y * w);
                                atomicAdd(sgrid + off + 2, gq4.z * w); atomicAdd(sgrid + off + 3, gq4.w * w);
                            } else {
                                red_add_v4(gg + off, gq4.x * w, gq4.y * w, gq4.z * w, gq4.w * w);
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();      // tiles / masks are rewritten by the next batch
    }

    // ---- flush: T fragments -> CTA accumulator in smem -> head gradients -> one RED per CTA per element
    __syncthreads();
    if constexpr (!SPARSE) {
        if (smem_grid_floats > 0 && grad.grids[0])
            for (int e = threadIdx.x * 4; e < smem_grid_floats; e += blockDim.x * 4) {
                const float4 v = *reinterpret_cast<const float4*>(sgrid + e);
                if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) red_add_v4(grad.grids[0] + e, v.x, v.y, v.z, v.w);
            }
    }
    float* cta_T = smem + SDF_SMEM_WARP_OFF;             // [128][40] (per-warp regions are dead now), then db1, loss
    for (int e = threadIdx.x; e < NGLOD_H * 40 + 4; e += blockDim.x) cta_T[e] = 0.f;
    __syncthreads();
#pragma unroll
    for (int nt = 0; nt < 5; ++nt) {
        const int h = 16 * mt + g, k = 8 * nt + 2 * t;
        atomicAdd(cta_T + h * 40 + k, accT[nt][0]);
        atomicAdd(cta_T + h * 40 + k + 1, accT[nt][1]);
        atomicAdd(cta_T + (h + 8) * 40 + k, accT[nt][2]);
        atomicAdd(cta_T + (h + 8) * 40 + k + 1, accT[nt][3]);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        acc_b1 += __shfl_xor_sync(0xffffffffu, acc_b1, o);
        acc_loss += __shfl_xor_sync(0xffffffffu, acc_loss, o);
    }
    if (lane == 0) {
        atomicAdd(cta_T + NGLOD_H * 40, acc_b1);
        atomicAdd(cta_T + NGLOD_H * 40 + 1, acc_loss);
    }
    __syncthreads();
    const int in_dim = net.pos_invariant ? NGLOD_F : NGLOD_F + 3;
    if (threadIdx.x < NGLOD_H) {
        const int h = threadIdx.x;
        const float w1h = sw1[h];
        float dw1 = 0.f;
#pragma unroll 4
        for (int k = 0; k < NGLOD_KPAD; ++k) {
            const float tv = cta_T[h * 40 + k];
            dw1 = fmaf(smem[h * NGLOD_KPAD + k] + wlo[h * NGLOD_KPAD + k], tv, dw1);   // W0ext row {32 feat, x, y, z, b0} = hi + lo
            const float v = w1h * tv;
            if (k < NGLOD_F) {
                if (grad.w0) atomicAdd(grad.w0 + h * in_dim + (net.pos_invariant ? k : k + 3), v);
            } else if (k < NGLOD_F + 3) {
                if (grad.w0 && !net.pos_invariant) atomicAdd(grad.w0 + h * in_dim + (k - NGLOD_F), v);
            } else {
                if (grad.b0) atomicAdd(grad.b0 + h, v);
            }
        }
        if (grad.w1) atomicAdd(grad.w1 + h, dw1);
    }
    if (threadIdx.x == 0) {
        if (grad.b1) atomicAdd(grad.b1, cta_T[NGLOD_H * 40]);
        if (FUSED_LOSS && loss_out) atomicAdd(loss_out, cta_T[NGLOD_H * 40 + 1]);
    }
}

// ---- transpose of the prefix sum: push dL/d(summed grid of level l) down to the LOD grids, level by level.
// restrict: Tc[c] += sum over fine nodes n in the support of coarse node c's hat function of  w(n, c) * Tf[n],
//           w = prod_axis (1 - |n_a - k c_a| / k), k = Rf / Rc  -- the weights nglod_build_summed_grid used, transposed.
__global__ void __launch_bounds__(256)
restrict_add_kernel(const float4* __restrict__ Tf, const int Rf, float4* __restrict__ Tc, const int Rc, const long long n_chunks) {
    const int Sc = Rc + 1, Sf = Rf + 1, k = Rf / Rc;
    const float inv_k = 1.f / (float)k;           // exact for the power-of-two ratios of an octree
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_chunks; e += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(e & 7);
        long long node = e >> 3;
        const int cx = (int)(node % Sc); node /= Sc;
        const int cy = (int)(node % Sc);
        const int cz = (int)(node / Sc);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
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
__global__ void __launch_bounds__(256)
fold_copies_kernel(float4* __restrict__ priv, const int copies, const long long n4, float4* __restrict__ target) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n4) return;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = 0; c < copies; ++c) {
        const float4 v = priv[c * n4 + e];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        priv[c * n4 + e] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (target) {
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
            restrict_add_kernel<<<(int)blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(grad->summed[l]), net->grid_res[l],
                                                             reinterpret_cast<float4*>(grad->summed[l - 1]), net->grid_res[l - 1], nc);
        }
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
    long long grid = nglod_sm_count();
    if constexpr (!WITH_GX) {
#ifndef NGLOD_BWD_TC
#define NGLOD_BWD_TC 1          // 0: the mma.sync kernel for the single-grid path too (A/B experiments)
#endif
#if NGLOD_BWD_TC
        // third generation (sdf_backward_tc.cu): tcgen05 GEMMs, warp-specialised; single-grid path
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
            if (int e = nglod_launch_sdf_backward_tc(nd, gdv, x, (long long)n, grad_out, gt, loss_scale, loss_out, FUSED_LOSS, st)) return e;
            if (gdv.priv) {
                const long long n4 = grid_floats / 4;
                fold_copies_kernel<<<(int)((n4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<float4*>(gdv.priv), gdv.priv_copies, n4,
                                                                            reinterpret_cast<float4*>(gdv.grids[0]));
                if (int e = (int)cudaGetLastError()) return e;
            }
            return cascade ? restrict_cascade(net, lod, grad, st) : 0;
        }
#endif
        // second-generation kernel: 16 warps, head gradients on mma.sync tensor cores (the first generation below serves dL/dx)
        {
            auto k2 = sdf_backward_mma_kernel<FUSED_LOSS, false>;
            NGLOD_CUDA_TRY(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, BW2_SMEM_BYTES + BW2_SMEM_GRID_MAX * 4));
            const int lpw = n > (long long)grid * BW2_WARPS * 8 ? 32 : 8;     // 128-query batches only while they fit one wave
            const long long want2 = (n + BW2_WARPS * lpw - 1) / (BW2_WARPS * lpw);
            if (want2 < grid) grid = want2;
            // level 0 of the launch small enough (R = 4) and busy enough to be worth a per-CTA copy in shared memory
            int sg = 0;
            {
                const long long nodes = (long long)(nd.res[0] + 1) * (nd.res[0] + 1) * (nd.res[0] + 1) * NGLOD_F;
                if (gdv.grids[0] && nodes <= BW2_SMEM_GRID_MAX && n >= 16 * nodes) sg = (int)nodes;
            }
            k2<<<(int)grid, BW2_THREADS, BW2_SMEM_BYTES + sg * 4, st>>>(nd, gdv, x, (long long)n, grad_out, gt, loss_scale, loss_out,
                                                                        SparseBwd{}, lpw, sg);
            if (int e = (int)cudaGetLastError()) return e;
            return (single && cascade) ? restrict_cascade(net, lod, grad, st) : 0;
        }
    }
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
    auto k2 = sdf_backward_mma_kernel<FUSED_LOSS, true>;
    NGLOD_CUDA_TRY(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, BW2_SMEM_BYTES));
    long long grid = nglod_sm_count();
    const int lpw = n > (long long)grid * BW2_WARPS * 8 ? 32 : 8;     // 128-query batches only while they fit one wave
    const long long want = (n + BW2_WARPS * lpw - 1) / (BW2_WARPS * lpw);
    if (want < grid) grid = want;
    k2<<<(int)grid, BW2_THREADS, BW2_SMEM_BYTES, (cudaStream_t)stream>>>(sp.sn.dec, gdv, x, (long long)n, grad_out, gt,
                                                                         loss_scale, loss_out, sp, lpw, 0);
    return (int)cudaGetLastError();
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
