// nglod_b200 -- sparse OctreeSDF tables as the kernels see them, the parent-chain gather (shared by the in-voxel tracer,
// the sparse forward and the sparse backward).  Behavioural spec: sparse_grid_sample_kernel,
// sol-renderer/include/solr/solr/sdf/sparse_grid_sample.cuh:31-109.
#pragma once
#include "common.cuh"

struct SparseDev {
    NetDev dec;                 // decoder of the selected LOD (grids unused)
    const float* cf;
    const int* trinkets;
    const int* parents;
    const short4* voxels;
    int lod;                    // LOD being evaluated
    int first_lod;              // first level to sample: 0 (walk the parent chain) or lod (cf = prefix-summed rows)
    int base_lod;
    int vox_off;                // first voxel row of `lod`
};

// Sparse backward: the features come from (and their gradients go to) the corner rows of a sparse octree model.
struct SparseBwd {
    SparseDev sn;
    const int* pidx;        // [n] voxel index within the LOD's level (< 0: point outside the octree, inert row)
    float* grad_cf;         // [NC, F]
};

// 4 channels [4c,4c+4) of sum_{l<=lod} trilinear(corner features) for a point in voxel row `vrow` of LOD sn.lod.
__device__ __forceinline__ float4 sparse_gather4(const SparseDev& sn, float qx, float qy, float qz, int vrow, int c) {
    int chain[NGLOD_MAX_LODS];
    {
        int v = vrow;
#pragma unroll
        for (int l = NGLOD_MAX_LODS - 1; l >= 0; --l) {
            if (l > sn.lod || l < sn.first_lod) continue;
            chain[l] = v;
            if (l > sn.first_lod) v = __ldg(sn.parents + v);
        }
    }
    const float nx = fmaf(qx, 0.5f, 0.5f), ny = fmaf(qy, 0.5f, 0.5f), nz = fmaf(qz, 0.5f, 0.5f);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int l = 0; l < NGLOD_MAX_LODS; ++l) {
        if (l > sn.lod) break;
        if (l < sn.first_lod) continue;
        const int v = chain[l];
        const float res = (float)(1 << (l + sn.base_lod));
        const short4 vc = __ldg(sn.voxels + v);
        const float fx = nx * res - (float)vc.x, fy = ny * res - (float)vc.y, fz = nz * res - (float)vc.z;
        const float gx = 1.f - fx, gy = 1.f - fy, gz = 1.f - fz;
        const int4 t0 = __ldg(reinterpret_cast<const int4*>(sn.trinkets + 8 * v));
        const int4 t1 = __ldg(reinterpret_cast<const int4*>(sn.trinkets + 8 * v) + 1);
        const float* base = sn.cf + 4 * c;
        float4 vv[8];
        vv[0] = ldg_f4(base + (size_t)t0.x * NGLOD_F); vv[1] = ldg_f4(base + (size_t)t0.y * NGLOD_F);
        vv[2] = ldg_f4(base + (size_t)t0.z * NGLOD_F); vv[3] = ldg_f4(base + (size_t)t0.w * NGLOD_F);
        vv[4] = ldg_f4(base + (size_t)t1.x * NGLOD_F); vv[5] = ldg_f4(base + (size_t)t1.y * NGLOD_F);
        vv[6] = ldg_f4(base + (size_t)t1.z * NGLOD_F); vv[7] = ldg_f4(base + (size_t)t1.w * NGLOD_F);
        const float w00 = gx * gy, w10 = fx * gy, w01 = gx * fy, w11 = fx * fy;
        const float w[8] = {w00 * gz, w10 * gz, w01 * gz, w11 * gz, w00 * fz, w10 * fz, w01 * fz, w11 * fz};
        float4 s;
        s.x = vv[0].x * w[0]; s.y = vv[0].y * w[0]; s.z = vv[0].z * w[0]; s.w = vv[0].w * w[0];
#pragma unroll
        for (int k = 1; k < 8; ++k) {
            s.x = fmaf(vv[k].x, w[k], s.x); s.y = fmaf(vv[k].y, w[k], s.y);
            s.z = fmaf(vv[k].z, w[k], s.z); s.w = fmaf(vv[k].w, w[k], s.w);
        }
        acc.x = s.x + acc.x; acc.y = s.y + acc.y; acc.z = s.z + acc.z; acc.w = s.w + acc.w;
    }
    return acc;
}

// Transpose of sparse_gather4: add w_{l,k} * g (channels [4c,4c+4) of dL/dfeat of one query) to the corner rows of every
// voxel of the query's parent chain.  grad_cf is shaped like the corner features.
__device__ __forceinline__ void sparse_scatter4(const SparseDev& sn, float* __restrict__ grad_cf, float qx, float qy, float qz,
                                                int vrow, int c, float4 g) {
    const float nx = fmaf(qx, 0.5f, 0.5f), ny = fmaf(qy, 0.5f, 0.5f), nz = fmaf(qz, 0.5f, 0.5f);
    int v = vrow;
#pragma unroll 1
    for (int l = sn.lod; l >= sn.first_lod; --l) {
        const float res = (float)(1 << (l + sn.base_lod));
        const short4 vc = __ldg(sn.voxels + v);
        const float fx = nx * res - (float)vc.x, fy = ny * res - (float)vc.y, fz = nz * res - (float)vc.z;
        const float gx = 1.f - fx, gy = 1.f - fy, gz = 1.f - fz;
        const int4 t0 = __ldg(reinterpret_cast<const int4*>(sn.trinkets + 8 * v));
        const int4 t1 = __ldg(reinterpret_cast<const int4*>(sn.trinkets + 8 * v) + 1);
        const int rows[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
        const float w00 = gx * gy, w10 = fx * gy, w01 = gx * fy, w11 = fx * fy;
        const float w[8] = {w00 * gz, w10 * gz, w01 * gz, w11 * gz, w00 * fz, w10 * fz, w01 * fz, w11 * fz};
#pragma unroll
        for (int k = 0; k < 8; ++k)
            red_add_v4(grad_cf + (size_t)rows[k] * NGLOD_F + 4 * c, g.x * w[k], g.y * w[k], g.z * w[k], g.w * w[k]);
        if (l > sn.first_lod) v = __ldg(sn.parents + v);
    }
}

static inline int make_sparse_dev(const nglod_sparse_net_t* net, int lod, SparseDev& sn, bool allow_summed = true) {
    if (!net) return NGLOD_EINVAL;
    if (net->num_lods < 1 || net->num_lods > NGLOD_MAX_LODS || lod < 0 || lod >= net->num_lods) return NGLOD_EINVAL;
    if (net->feature_dim != NGLOD_F || net->hidden_dim != NGLOD_H) return NGLOD_EUNSUPPORTED;
    if (net->math_mode != NGLOD_MATH_TC3XTF32 && net->math_mode != NGLOD_MATH_FP32) return NGLOD_EINVAL;
    if (!net->corner_feats || !net->trinkets || !net->parents || !net->voxels) return NGLOD_EINVAL;
    if ((reinterpret_cast<uintptr_t>(net->corner_feats) & 15u) || (reinterpret_cast<uintptr_t>(net->trinkets) & 15u) ||
        (reinterpret_cast<uintptr_t>(net->voxels) & 7u)) return NGLOD_EINVAL;
    if (!net->w0[lod] || !net->b0[lod] || !net->w1[lod] || !net->b1[lod]) return NGLOD_EINVAL;
    if (net->base_lod < 0 || net->base_lod + lod > 14) return NGLOD_EINVAL;
    sn.dec.num_lods = 0; sn.dec.pos_invariant = net->pos_invariant ? 1 : 0; sn.dec.half_pairs = 0;
    for (int i = 0; i < NGLOD_MAX_LODS; ++i) { sn.dec.res[i] = 1; sn.dec.grids[i] = nullptr; }
    sn.dec.w0 = net->w0[lod]; sn.dec.b0 = net->b0[lod]; sn.dec.w1 = net->w1[lod]; sn.dec.b1 = net->b1[lod];
    if (reinterpret_cast<uintptr_t>(net->corner_feats_summed) & 15u) return NGLOD_EINVAL;
    const bool summed = allow_summed && net->corner_feats_summed;
    sn.cf = summed ? net->corner_feats_summed : net->corner_feats;
    sn.first_lod = summed ? lod : 0;
    sn.trinkets = net->trinkets; sn.parents = net->parents;
    sn.voxels = reinterpret_cast<const short4*>(net->voxels);
    sn.lod = lod; sn.base_lod = net->base_lod; sn.vox_off = net->lod_voxel_offset[lod];
    return 0;
}

