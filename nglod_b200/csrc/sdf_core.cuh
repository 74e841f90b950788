// nglod_b200 -- warp-cooperative OctreeSDF evaluation (FP32 "exact" path).
//
// What it computes (reference: sdf-net/lib/models/OctreeSDF.py:46-57,94-146):
//   feat = sum_{i<=lod} trilinear(grid_i, p)        (F.grid_sample, align_corners, border)
//   d    = W1 . relu(W0 . [p, feat] + b0) + b1
//
// How (B200-first, not the reference's per-LOD launch chain):
//  * Grids are channels-last, so one corner = 32 fp32 = ONE 128-byte line.
//    A sub-warp of 8 lanes owns one query; lane c loads channels 4c..4c+3 of
//    every corner with one LDG.128, so each warp-level load instruction moves
//    4 full lines (4 L1 wavefronts) instead of 32 scattered sectors.
//  * Interpolated features never leave the SM: they go registers -> a 4.6 KB
//    per-warp shared tile -> registers of the lane that owns the query.
//  * The 35->128->1 decoder runs thread-per-query with W0 (bias folded in as a
//    36th column against a constant-1 input) broadcast from shared memory.
//  * The same routine is called inline by the sphere tracer with an arbitrary
//    subset of lanes active; live lanes are compacted with ballot/popc so the
//    gather rounds only run for live queries.
#pragma once
#include "common.cuh"

#define SDF_WARPS 8                    // warps per CTA for every SDF-evaluating kernel
#define SDF_THREADS (SDF_WARPS * 32)
#define SDF_W0_FLOATS (NGLOD_H * NGLOD_KPAD)
#define SDF_TILE_FLOATS (32 * NGLOD_KPAD)

// dynamic shared memory layout (floats):
//   [0, 4608)            W0 permuted: row j = {w_feat[0..31], w_x, w_y, w_z, b0[j]}
//   [4608, 4736)         W1
//   [4736, 4740)         b1 (+pad)
//   then per warp:       tile[32][36]  (row q = {feat[0..31], x, y, z, 1})
//                        idx[32] (int) : slot -> lane of the slot-th live query
#define SDF_SMEM_W1_OFF (SDF_W0_FLOATS)
#define SDF_SMEM_B1_OFF (SDF_W0_FLOATS + NGLOD_H)
#define SDF_SMEM_WARP_OFF (SDF_W0_FLOATS + NGLOD_H + 4)
#define SDF_SMEM_PER_WARP (SDF_TILE_FLOATS + 32)
#define SDF_SMEM_BYTES ((SDF_SMEM_WARP_OFF + SDF_WARPS * SDF_SMEM_PER_WARP) * 4)

// Stage one decoder head into shared memory (all threads of the CTA).
__device__ __forceinline__ void sdf_stage_weights(const NetDev& net, float* smem) {
    const int in_dim = net.pos_invariant ? NGLOD_F : NGLOD_F + 3;
    // batches of 6 independent loads, then 6 stores: with a cold L2 a load-store loop pays the DRAM latency per iteration
    for (int base = 0; base < SDF_W0_FLOATS; base += 6 * blockDim.x) {
        float v[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const int e = base + threadIdx.x + i * blockDim.x;
            const int j = e / NGLOD_KPAD, k = e - j * NGLOD_KPAD;
            v[i] = 0.f;
            if (e < SDF_W0_FLOATS) {
                if (k < NGLOD_F) v[i] = __ldg(net.w0 + j * in_dim + (net.pos_invariant ? k : k + 3));
                else if (k < NGLOD_F + 3) v[i] = net.pos_invariant ? 0.f : __ldg(net.w0 + j * in_dim + (k - NGLOD_F));
                else v[i] = __ldg(net.b0 + j);
            }
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const int e = base + threadIdx.x + i * blockDim.x;
            if (e < SDF_W0_FLOATS) smem[e] = v[i];
        }
    }
    for (int e = threadIdx.x; e < NGLOD_H; e += blockDim.x) smem[SDF_SMEM_W1_OFF + e] = __ldg(net.w1 + e);
    if (threadIdx.x == 0) smem[SDF_SMEM_B1_OFF] = __ldg(net.b1);
}

// Trilinear set-up for one LOD, exactly PyTorch's grid_sampler_3d arithmetic
// (aten/src/ATen/native/GridSampler.h: unnormalize with align_corners, clip to
// the border, floor, weights from (corner+1 - u) and (u - corner)).
struct LodSetup {
    int off[8];     // element offsets of the 8 corners (x fastest, then y, then z)
    float w[8];
};

__device__ __forceinline__ void lod_axis(float p, int R, int& i0, int& i1, float& w0, float& w1) {
    const float fR = (float)R;
    float u = ((p + 1.f) * 0.5f) * fR;
    u = fminf(fR, fmaxf(u, 0.f));
    const float f0 = floorf(u);
    i0 = (int)f0;
    i1 = min(i0 + 1, R);            // out-of-range corner has weight exactly 0
    w0 = (f0 + 1.f) - u;
    w1 = u - f0;
}

__device__ __forceinline__ void lod_setup(float px, float py, float pz, int R, LodSetup& s) {
    int x0, x1, y0, y1, z0, z1;
    float wx0, wx1, wy0, wy1, wz0, wz1;
    lod_axis(px, R, x0, x1, wx0, wx1);
    lod_axis(py, R, y0, y1, wy0, wy1);
    lod_axis(pz, R, z0, z1, wz0, wz1);
    const int S = R + 1;
    const int zy00 = (z0 * S + y0) * S, zy01 = (z0 * S + y1) * S;
    const int zy10 = (z1 * S + y0) * S, zy11 = (z1 * S + y1) * S;
    s.off[0] = (zy00 + x0) * NGLOD_F; s.w[0] = (wx0 * wy0) * wz0;
    s.off[1] = (zy00 + x1) * NGLOD_F; s.w[1] = (wx1 * wy0) * wz0;
    s.off[2] = (zy01 + x0) * NGLOD_F; s.w[2] = (wx0 * wy1) * wz0;
    s.off[3] = (zy01 + x1) * NGLOD_F; s.w[3] = (wx1 * wy1) * wz0;
    s.off[4] = (zy10 + x0) * NGLOD_F; s.w[4] = (wx0 * wy0) * wz1;
    s.off[5] = (zy10 + x1) * NGLOD_F; s.w[5] = (wx1 * wy0) * wz1;
    s.off[6] = (zy11 + x0) * NGLOD_F; s.w[6] = (wx0 * wy1) * wz1;
    s.off[7] = (zy11 + x1) * NGLOD_F; s.w[7] = (wx1 * wy1) * wz1;
}

// Sum over LODs of the trilinear sample, for the 4 channels [4c, 4c+4) of one query.
__device__ __forceinline__ float4 gather_features(const NetDev& net, float qx, float qy, float qz, int c) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < NGLOD_MAX_LODS; ++i) {
        if (i >= net.num_lods) break;
        LodSetup s;
        lod_setup(qx, qy, qz, net.res[i], s);
        const float* g = net.grids[i] + 4 * c;
        float4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = ldg_f4(g + s.off[k]);
        float4 r;
        r.x = v[0].x * s.w[0]; r.y = v[0].y * s.w[0]; r.z = v[0].z * s.w[0]; r.w = v[0].w * s.w[0];
#pragma unroll
        for (int k = 1; k < 8; ++k) {
            r.x = fmaf(v[k].x, s.w[k], r.x); r.y = fmaf(v[k].y, s.w[k], r.y);
            r.z = fmaf(v[k].z, s.w[k], r.z); r.w = fmaf(v[k].w, s.w[k], r.w);
        }
        // running sum across LODs (OctreeSDF.py:109-110)
        acc.x = r.x + acc.x; acc.y = r.y + acc.y; acc.z = r.z + acc.z; acc.w = r.w + acc.w;
    }
    return acc;
}

// Warp-cooperative feature gather into the warp's shared tile.
// Every lane of the warp must call this (convergent).  On return (after the
// trailing __syncwarp) tile row `lane` holds {feat[32], x, y, z, 1} for each
// active lane.
__device__ __forceinline__ void warp_gather_tile(const NetDev& net, float px, float py, float pz,
                                                 bool active, float* tile, int* idx, int lane) {
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
        const float qx = __shfl_sync(0xffffffffu, px, q);
        const float qy = __shfl_sync(0xffffffffu, py, q);
        const float qz = __shfl_sync(0xffffffffu, pz, q);
        if (valid) {
            const float4 f = gather_features(net, qx, qy, qz, c);
            *reinterpret_cast<float4*>(tile + q * NGLOD_KPAD + 4 * c) = f;
        }
    }
    __syncwarp();
}

// Decoder for the calling lane's own tile row.  Returns d.  `hidden_out`, when
// non-null, receives nothing here (the backward kernel has its own variant).
__device__ __forceinline__ float lane_decoder(const float* __restrict__ sW, const float* __restrict__ row) {
    float in[NGLOD_KPAD];
#pragma unroll
    for (int k4 = 0; k4 < NGLOD_KPAD / 4; ++k4) {
        const float4 v = *reinterpret_cast<const float4*>(row + 4 * k4);
        in[4 * k4] = v.x; in[4 * k4 + 1] = v.y; in[4 * k4 + 2] = v.z; in[4 * k4 + 3] = v.w;
    }
    const float4* w4 = reinterpret_cast<const float4*>(sW);
    const float* w1 = sW + SDF_SMEM_W1_OFF;
    float out = sW[SDF_SMEM_B1_OFF];
#pragma unroll 1
    for (int j = 0; j < NGLOD_H; j += 4) {
        float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k4 = 0; k4 < NGLOD_KPAD / 4; ++k4) {
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const float4 w = w4[(j + jj) * (NGLOD_KPAD / 4) + k4];
                a[jj] = fmaf(w.x, in[4 * k4], a[jj]);
                a[jj] = fmaf(w.y, in[4 * k4 + 1], a[jj]);
                a[jj] = fmaf(w.z, in[4 * k4 + 2], a[jj]);
                a[jj] = fmaf(w.w, in[4 * k4 + 3], a[jj]);
            }
        }
        const float4 v1 = *reinterpret_cast<const float4*>(w1 + j);
        out = fmaf(v1.x, fmaxf(a[0], 0.f), out);
        out = fmaf(v1.y, fmaxf(a[1], 0.f), out);
        out = fmaf(v1.z, fmaxf(a[2], 0.f), out);
        out = fmaf(v1.w, fmaxf(a[3], 0.f), out);
    }
    return out;
}

// Full evaluation: every lane of the warp calls; returns sdf(p) for active lanes.
__device__ __forceinline__ float warp_sdf_eval(const NetDev& net, const float* sW, float* tile, int* idx,
                                               float px, float py, float pz, bool active, int lane) {
    warp_gather_tile(net, px, py, pz, active, tile, idx, lane);
    float d = 0.f;
    if (__any_sync(0xffffffffu, active)) {
        // inactive lanes run on stale rows; their result is discarded
        d = lane_decoder(sW, tile + lane * NGLOD_KPAD);
    }
    __syncwarp();
    return d;
}
