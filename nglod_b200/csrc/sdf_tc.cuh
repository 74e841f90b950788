// nglod_b200 -- tensor-core OctreeSDF evaluation: building blocks shared by the forward kernel and the tracer.
//
// A "group" is 4 consecutive warps (128 threads) that own one 128-row MMA tile: every lane owns one query = one
// A row = one TMEM lane.  Per tile:
//   1. tc_gather_rows   features (8 lanes per corner line, per-LOD set-up computed ONCE per query and shared through
//                       a 16-byte smem record) -> split hi/lo TF32 -> A_hi / A_lo rows in the UMMA smem layout
//   2. fence.proxy.async + group barrier; one thread issues 15 tcgen05.mma (3xTF32) + tcgen05.commit -> mbarrier
//   3. tc_epilogue      each lane reads its 128 accumulator columns from TMEM, d = b1 + sum_j W1[j]*relu(D[j])
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

#define TCG_THREADS 128                 // threads per group
#define TC_PACK_LODS 5                  // LODs whose set-up records fit the per-warp scratch at once
#define TC_WARP_SCRATCH_BYTES (TC_PACK_LODS * 32 * 16 + 32 * 4)   // records + slot->lane index

// smem carve-up for a CTA with G groups
#define TC_SMEM_B_HI 0
#define TC_SMEM_B_LO (TC_OPERAND_BYTES)
#define TC_SMEM_A(g) (2 * TC_OPERAND_BYTES + (g) * 2 * TC_OPERAND_BYTES)          // A_hi of group g; A_lo follows
#define TC_SMEM_W1(G) (2 * TC_OPERAND_BYTES + (G) * 2 * TC_OPERAND_BYTES)           // 128 floats + b1 (+pad) = 528 B
#define TC_SMEM_SCRATCH(G) (TC_SMEM_W1(G) + 528)
#define TC_SMEM_MBAR(G) (TC_SMEM_SCRATCH(G) + (G) * 4 * TC_WARP_SCRATCH_BYTES)      // G mbarriers (8 B each)
#define TC_SMEM_TMEMPTR(G) (TC_SMEM_MBAR(G) + 8 * (G))
#define TC_SMEM_BYTES(G) (TC_SMEM_TMEMPTR(G) + 16)

// Stage W0|b0 as the B operand (hi and lo), W1 and b1.  All threads of the CTA; followed by a __syncthreads by the caller.
__device__ __forceinline__ void tc_stage_weights(const NetDev& net, char* smem, int G) {
    const int in_dim = net.pos_invariant ? NGLOD_F : NGLOD_F + 3;
    for (int e = threadIdx.x; e < NGLOD_H * TC_K; e += blockDim.x) {
        const int j = e / TC_K, k = e - j * TC_K;
        float v = 0.f;
        if (k < NGLOD_F) v = __ldg(net.w0 + j * in_dim + (net.pos_invariant ? k : k + 3));
        else if (k < NGLOD_F + 3) v = net.pos_invariant ? 0.f : __ldg(net.w0 + j * in_dim + (k - NGLOD_F));
        else if (k == NGLOD_F + 3) v = __ldg(net.b0 + j);
        const float hi = tf32_hi(v);
        const uint32_t off = tc_elem_offset(j, k);
        *reinterpret_cast<float*>(smem + TC_SMEM_B_HI + off) = hi;
        *reinterpret_cast<float*>(smem + TC_SMEM_B_LO + off) = v - hi;
    }
    float* w1 = reinterpret_cast<float*>(smem + TC_SMEM_W1(G));
    for (int e = threadIdx.x; e < NGLOD_H; e += blockDim.x) w1[e] = __ldg(net.w1 + e);
    if (threadIdx.x == 0) w1[NGLOD_H] = __ldg(net.b1);
}

__device__ __forceinline__ void tc_store_split4(char* a_hi, char* a_lo, uint32_t off, float4 v) {
    float4 h, l;
    h.x = tf32_hi(v.x); h.y = tf32_hi(v.y); h.z = tf32_hi(v.z); h.w = tf32_hi(v.w);
    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
    *reinterpret_cast<float4*>(a_hi + off) = h;
    *reinterpret_cast<float4*>(a_lo + off) = l;
}

// One axis of the trilinear set-up (PyTorch grid_sampler arithmetic, see sdf_core.cuh::lod_axis).
// Returns floor index, the upper weight w1 = u - floor(u) and whether the +1 corner exists.
__device__ __forceinline__ void tc_axis(float p, int R, int& i0, float& w1, bool& has1) {
    const float fR = (float)R;
    float u = ((p + 1.f) * 0.5f) * fR;
    u = fminf(fR, fmaxf(u, 0.f));
    const float f0 = floorf(u);
    i0 = (int)f0;
    w1 = u - f0;
    has1 = i0 < R;
}

// Sum of NL LODs' trilinear samples for channels [4c,4c+4) of the query whose set-up records are pack[l*32+q].
template <int NL>
__device__ __forceinline__ float4 tc_gather_one(const NetDev& net, int l0, const float4* pack, int q, int c, float4 acc) {
    float4 P[NL];
#pragma unroll
    for (int l = 0; l < NL; ++l) P[l] = pack[l * 32 + q];
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        const uint32_t pk = __float_as_uint(P[l].x);
        const int S = net.res[l0 + l] + 1;
        const int o0 = (int)(pk & ~31u) + 4 * c;
        const int dx = (pk & 1u) ? NGLOD_F : 0;
        const int dy = (pk & 2u) ? S * NGLOD_F : 0;
        const int dz = (pk & 4u) ? S * S * NGLOD_F : 0;
        const float wx1 = P[l].y, wy1 = P[l].z, wz1 = P[l].w;
        const float wx0 = 1.f - wx1, wy0 = 1.f - wy1, wz0 = 1.f - wz1;   // == (floor+1) - u exactly
        const float* g = net.grids[l0 + l];
        const int o2 = o0 + dy, o4 = o0 + dz, o6 = o4 + dy;
        float4 v[8];
        v[0] = ldg_f4(g + o0); v[1] = ldg_f4(g + o0 + dx);
        v[2] = ldg_f4(g + o2); v[3] = ldg_f4(g + o2 + dx);
        v[4] = ldg_f4(g + o4); v[5] = ldg_f4(g + o4 + dx);
        v[6] = ldg_f4(g + o6); v[7] = ldg_f4(g + o6 + dx);
        const float w00 = wx0 * wy0, w10 = wx1 * wy0, w01 = wx0 * wy1, w11 = wx1 * wy1;
        const float w[8] = {w00 * wz0, w10 * wz0, w01 * wz0, w11 * wz0, w00 * wz1, w10 * wz1, w01 * wz1, w11 * wz1};
        float4 s;
        s.x = v[0].x * w[0]; s.y = v[0].y * w[0]; s.z = v[0].z * w[0]; s.w = v[0].w * w[0];
#pragma unroll
        for (int k = 1; k < 8; ++k) {
            s.x = fmaf(v[k].x, w[k], s.x); s.y = fmaf(v[k].y, w[k], s.y);
            s.z = fmaf(v[k].z, w[k], s.z); s.w = fmaf(v[k].w, w[k], s.w);
        }
        // running sum across LODs (OctreeSDF.py:109-110)
        acc.x = s.x + acc.x; acc.y = s.y + acc.y; acc.z = s.z + acc.z; acc.w = s.w + acc.w;
    }
    return acc;
}

template <int NL>
__device__ __forceinline__ void tc_gather_rounds(const NetDev& net, int l0, int n_live, char* a_hi, char* a_lo,
                                                 int row0, const float4* pack, const int* idx, int lane) {
    const int sub = lane >> 3, c = lane & 7;
    for (int r = 0; r * 4 < n_live; ++r) {
        const int slot = r * 4 + sub;
        if (slot < n_live) {
            const int q = idx[slot];
            const uint32_t row_off = tc_elem_offset(row0 + q, 4 * c);
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (l0 > 0) {       // more than TC_PACK_LODS grids: continue the running sum (hi + lo is exact)
                const float4 h = *reinterpret_cast<const float4*>(a_hi + row_off);
                const float4 lo = *reinterpret_cast<const float4*>(a_lo + row_off);
                acc = make_float4(h.x + lo.x, h.y + lo.y, h.z + lo.z, h.w + lo.w);
            }
            acc = tc_gather_one<NL>(net, l0, pack, q, c, acc);
            tc_store_split4(a_hi, a_lo, row_off, acc);
        }
    }
}

// Gather for the warp's 32 queries into rows [row0, row0+32) of the group's A operand.
//   px,py,pz / active : this lane's query;   pack/idx : this warp's scratch.
// Every lane of the warp must call (convergent).
__device__ __forceinline__ void tc_gather_rows(const NetDev& net, float px, float py, float pz, bool active,
                                               char* a_hi, char* a_lo, int row0, float4* pack, int* idx, int lane) {
    const unsigned live = __ballot_sync(0xffffffffu, active);
    const int n_live = __popc(live);
    if (n_live == 0) return;
    if (active) {
        idx[__popc(live & ((1u << lane) - 1u))] = lane;
        // K chunk 8 = {x, y, z, 1}: the query's own lane writes it (no shuffles needed later)
        tc_store_split4(a_hi, a_lo, tc_elem_offset(row0 + lane, NGLOD_F), make_float4(px, py, pz, 1.f));
    }
    for (int l0 = 0; l0 < net.num_lods; l0 += TC_PACK_LODS) {
        const int nl = min(TC_PACK_LODS, net.num_lods - l0);
        // ---- phase 1: per-LOD set-up, once per query (not once per lane of the query)
        if (active) {
            for (int l = 0; l < nl; ++l) {
                const int R = net.res[l0 + l], S = R + 1;
                int x0, y0, z0; float wx, wy, wz; bool hx, hy, hz;
                tc_axis(px, R, x0, wx, hx);
                tc_axis(py, R, y0, wy, hy);
                tc_axis(pz, R, z0, wz, hz);
                const uint32_t off = (uint32_t)(((z0 * S + y0) * S + x0) * NGLOD_F) | (hx ? 1u : 0u) | (hy ? 2u : 0u) | (hz ? 4u : 0u);
                pack[l * 32 + lane] = make_float4(__uint_as_float(off), wx, wy, wz);
            }
        }
        __syncwarp();
        // ---- phase 2: 4 queries per round, 8 lanes per corner line (LOD count compile-time so the loads of
        //      several LODs can be in flight together)
        switch (nl) {
            case 1: tc_gather_rounds<1>(net, l0, n_live, a_hi, a_lo, row0, pack, idx, lane); break;
            case 2: tc_gather_rounds<2>(net, l0, n_live, a_hi, a_lo, row0, pack, idx, lane); break;
            case 3: tc_gather_rounds<3>(net, l0, n_live, a_hi, a_lo, row0, pack, idx, lane); break;
            case 4: tc_gather_rounds<4>(net, l0, n_live, a_hi, a_lo, row0, pack, idx, lane); break;
            default: tc_gather_rounds<5>(net, l0, n_live, a_hi, a_lo, row0, pack, idx, lane); break;
        }
        __syncwarp();
    }
}

// d = b1 + sum_j W1[j] * relu(D[row, j]) for this lane's accumulator row (TMEM lane = 32*(warp%4) + lane).
__device__ __forceinline__ float tc_epilogue(uint32_t taddr_row, const float* __restrict__ w1) {
    float d = w1[NGLOD_H];
#pragma unroll
    for (int cb = 0; cb < NGLOD_H / 32; ++cb) {
        float v[32];
        tmem_ld32(taddr_row + cb * 32, v);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
            const float4 w = *reinterpret_cast<const float4*>(w1 + cb * 32 + 4 * j4);
            d = fmaf(w.x, fmaxf(v[4 * j4], 0.f), d);
            d = fmaf(w.y, fmaxf(v[4 * j4 + 1], 0.f), d);
            d = fmaf(w.z, fmaxf(v[4 * j4 + 2], 0.f), d);
            d = fmaf(w.w, fmaxf(v[4 * j4 + 3], 0.f), d);
        }
    }
    return d;
}

// Per-group context + one full tile evaluation: gather -> MMA -> epilogue.  All 128 threads of the group call.
struct TcGroup {
    char* a_hi; char* a_lo;
    uint32_t a_hi_s, a_lo_s, b_hi_s, b_lo_s;   // shared-space addresses for descriptors
    uint32_t mbar_s;
    uint32_t tmem_row;                          // TMEM address of this lane's row, column 0 of the group's accumulator
    uint32_t tmem_acc;                          // TMEM address of the accumulator (lane 0)
    float4* pack; int* idx;
    const float* w1;
    int wq, lane, bar_id;
    uint32_t parity;
};

__device__ __forceinline__ float tc_group_eval(const NetDev& net, TcGroup& g, float px, float py, float pz, bool active) {
    tc_gather_rows(net, px, py, pz, active, g.a_hi, g.a_lo, g.wq * 32, g.pack, g.idx, g.lane);
    fence_proxy_async_smem();                     // generic-proxy smem writes -> visible to the tensor core
    tc_fence_before_sync();                       // order the previous tile's TMEM loads before the next MMA
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

__device__ __forceinline__ TcGroup tc_make_group(char* smem, int G, uint32_t tmem_base) {
    TcGroup g;
    const int warp = threadIdx.x >> 5;
    const int grp = warp >> 2;
    g.wq = warp & 3;
    g.lane = threadIdx.x & 31;
    g.a_hi = smem + TC_SMEM_A(grp);
    g.a_lo = g.a_hi + TC_OPERAND_BYTES;
    g.a_hi_s = smem_u32(g.a_hi);
    g.a_lo_s = smem_u32(g.a_lo);
    g.b_hi_s = smem_u32(smem + TC_SMEM_B_HI);
    g.b_lo_s = smem_u32(smem + TC_SMEM_B_LO);
    g.mbar_s = smem_u32(smem + TC_SMEM_MBAR(G) + 8 * grp);
    g.tmem_acc = tmem_base + (uint32_t)(grp * TC_N);
    g.tmem_row = g.tmem_acc + ((uint32_t)(g.wq * 32) << 16);
    char* scratch = smem + TC_SMEM_SCRATCH(G) + warp * TC_WARP_SCRATCH_BYTES;
    g.pack = reinterpret_cast<float4*>(scratch);
    g.idx = reinterpret_cast<int*>(scratch + TC_PACK_LODS * 32 * 16);
    g.w1 = reinterpret_cast<const float*>(smem + TC_SMEM_W1(G));
    g.bar_id = 1 + grp;
    g.parity = 0;
    return g;
}

// Common prologue: zero the operand buffers, stage weights, init mbarriers, allocate TMEM.  Returns the TMEM base.
__device__ __forceinline__ uint32_t tc_prologue(const NetDev& net, char* smem, int G) {
    for (int e = threadIdx.x; e < (TC_SMEM_SCRATCH(G) + G * 4 * TC_WARP_SCRATCH_BYTES) / 16; e += blockDim.x)
        reinterpret_cast<float4*>(smem)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    tc_stage_weights(net, smem, G);
    if (threadIdx.x == 0) {
        for (int g = 0; g < G; ++g) mbar_init(smem_u32(smem + TC_SMEM_MBAR(G) + 8 * g), 1);
        mbar_fence_init();
    }
    if (threadIdx.x < 32) tmem_alloc(smem_u32(smem + TC_SMEM_TMEMPTR(G)), 512);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    return *reinterpret_cast<volatile uint32_t*>(smem + TC_SMEM_TMEMPTR(G));
}

__device__ __forceinline__ void tc_epilogue_free(uint32_t tmem_base) {
    tc_fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem_base, 512);
}


// OR-reduce a predicate over the 128 threads of a group (also a barrier).
__device__ __forceinline__ bool tc_group_any(int bar_id, bool pred) {
    uint32_t r;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
        "bar.red.or.pred p, %2, %3, q;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(r) : "r"((uint32_t)pred), "r"(bar_id), "r"(TCG_THREADS) : "memory");
    return r != 0;
}
