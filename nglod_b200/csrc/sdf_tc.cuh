// nglod_b200 -- tensor-core OctreeSDF evaluation: building blocks shared by the forward kernel and the tracer.
//
// A "group" is 4 consecutive warps (128 threads) that own one 128-row MMA tile: every lane owns one query = one
// A row = one TMEM lane.  Per tile:
//   1. tc_gather_rows   features, 8 lanes per 128-byte line, FFMA2 interpolation -> split hi/lo TF32 -> A_hi / A_lo rows
//                       in the UMMA smem layout.  Three kernel flavours (TC_MULTI / TC_SINGLE_F32 / TC_SINGLE_HALF):
//                       per-LOD fp32 grids (set-up records staged in smem once per query), or ONE prefix-summed grid
//                       in fp32 corner lines or fp16 x-pair lines (set-up records travel by warp shuffle, no scratch)
//   2. fence.proxy.async + group barrier (bar.red.or: doubles as the "anyone still working?" vote of the persistent
//      tracer); one thread issues 15 tcgen05.mma (3xTF32) + tcgen05.commit -> mbarrier            [tc_group_issue]
//   3. tc_epilogue      each lane reads its 128 accumulator columns from TMEM, d = b1 + sum_j W1[j]*relu(D[j])
//                                                                                                  [tc_group_finish]
#pragma once
#include "common.cuh"
#include "tc_common.cuh"
#include <cuda_fp16.h>

#define TCG_THREADS 128                 // threads per group
#define TC_PACK_LODS 5                  // LODs whose set-up records fit the per-warp scratch at once
#define TC_WARP_SCRATCH_BYTES (TC_PACK_LODS * 32 * 16 + 32 * 4)   // records + slot->lane index

// smem carve-up for a CTA with G groups
#define TC_SMEM_B_HI 0
#define TC_SMEM_B_LO (TC_OPERAND_BYTES)
#define TC_SMEM_A(g) (2 * TC_OPERAND_BYTES + (g) * 2 * TC_OPERAND_BYTES)          // A_hi of group g; A_lo follows
#define TC_SMEM_W1(G) (2 * TC_OPERAND_BYTES + (G) * 2 * TC_OPERAND_BYTES)           // 128 floats + b1 (+pad) = 528 B
#define TC_SMEM_SCRATCH(G) (TC_SMEM_W1(G) + 528)
// W = per-warp scratch bytes: TC_WARP_SCRATCH_BYTES for kernels that stage set-up records (multi-LOD gather, sparse
// tracer), 0 for the single-grid kernels (records travel by warp shuffle)
#define TC_SMEM_MBAR_W(G, W) (TC_SMEM_SCRATCH(G) + (G) * 4 * (W))                   // G mbarriers (8 B each)
#define TC_SMEM_TMEMPTR_W(G, W) (TC_SMEM_MBAR_W(G, W) + 8 * (G))
#define TC_SMEM_BYTES_W(G, W) (TC_SMEM_TMEMPTR_W(G, W) + 16)
#define TC_SMEM_BYTES(G) TC_SMEM_BYTES_W(G, TC_WARP_SCRATCH_BYTES)

// which gather a kernel instance is built for (separate instances keep register allocation clean)
enum : int { TC_MULTI = 0, TC_SINGLE_F32 = 1, TC_SINGLE_HALF = 2 };
__host__ __device__ constexpr int tc_mode_scratch(int mode) { return mode == TC_MULTI ? TC_WARP_SCRATCH_BYTES : 0; }

// Stage W0|b0 as the B operand (hi and lo), W1 and b1.  All threads of the CTA.  The global loads are issued as ONE
// batch of independent coalesced loads before anything is written: with a cold L2 (first launch after the weights
// changed, or a flushed cache) a load-store-load-store loop pays the DRAM latency once per iteration (~20 us per
// launch, measured), the batch pays it once.
__device__ __forceinline__ void tc_load_weights(const NetDev& net, int base, float (&v)[16]) {
    const int in_dim = net.pos_invariant ? NGLOD_F : NGLOD_F + 3;
    const int n_w0 = NGLOD_H * in_dim;
    const int total = n_w0 + 2 * NGLOD_H + 1;              // W0, b0, W1, b1 (each its own pointer)
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int e = base + threadIdx.x + i * blockDim.x;
        float x = 0.f;
        if (e < n_w0) x = __ldg(net.w0 + e);
        else if (e < n_w0 + NGLOD_H) x = __ldg(net.b0 + (e - n_w0));
        else if (e < n_w0 + 2 * NGLOD_H) x = __ldg(net.w1 + (e - n_w0 - NGLOD_H));
        else if (e < total) x = __ldg(net.b1);
        v[i] = x;
    }
}
__device__ __forceinline__ void tc_scatter_weights_to(const NetDev& net, char* smem, float* w1, int base, const float (&v)[16]) {
    const int in_dim = net.pos_invariant ? NGLOD_F : NGLOD_F + 3;
    const int n_w0 = NGLOD_H * in_dim;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int e = base + threadIdx.x + i * blockDim.x;
        if (e < n_w0 + NGLOD_H) {
            int j, k;                                        // B row (hidden unit), K column
            if (e < n_w0) {
                j = e / in_dim;
                const int c = e - j * in_dim;                // torch column: xyz first (unless pos_invariant)
                k = net.pos_invariant ? c : (c < 3 ? NGLOD_F + c : c - 3);
            } else {
                j = e - n_w0;
                k = NGLOD_F + 3;                             // bias rides on the constant-1 column
            }
            const float hi = tf32_hi(v[i]);
            const uint32_t off = tc_elem_offset(j, k);
            *reinterpret_cast<float*>(smem + TC_SMEM_B_HI + off) = hi;
            *reinterpret_cast<float*>(smem + TC_SMEM_B_LO + off) = v[i] - hi;
        } else if (e < n_w0 + 2 * NGLOD_H + 1) {
            w1[e - n_w0 - NGLOD_H] = v[i];                   // W1[0..127], then b1 at index 128
        }
    }
}

__device__ __forceinline__ void tc_store_split4(char* a_hi, char* a_lo, uint32_t off, float4 v) {
    float4 h, l;
    h.x = tf32_hi(v.x); h.y = tf32_hi(v.y); h.z = tf32_hi(v.z); h.w = tf32_hi(v.w);
    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
    *reinterpret_cast<float4*>(a_hi + off) = h;
    *reinterpret_cast<float4*>(a_lo + off) = l;
}

// Same values for finite v (see tf32_hi_finite): the warp-specialised kernel's version of tc_store_split4.
__device__ __forceinline__ void tc_store_split4_finite(char* a_hi, char* a_lo, uint32_t off, float4 v) {
    float4 h, l;
    h.x = tf32_hi_finite(v.x); h.y = tf32_hi_finite(v.y); h.z = tf32_hi_finite(v.z); h.w = tf32_hi_finite(v.w);
    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
    *reinterpret_cast<float4*>(a_hi + off) = h;
    *reinterpret_cast<float4*>(a_lo + off) = l;
}

__device__ __forceinline__ void tc_scatter_weights(const NetDev& net, char* smem, int G, int base, const float (&v)[16]) {
    tc_scatter_weights_to(net, smem, reinterpret_cast<float*>(smem + TC_SMEM_W1(G)), base, v);
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

// two fp16 values in one 32-bit word -> packed fp32x2 (exact)
__device__ __forceinline__ uint64_t h2_to_f2(uint32_t h) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h));
    return f2_pack(f.x, f.y);
}

// ------------------------------------------------------------------------------------------------ gather
// Set-up record of one (query, LOD): {packed offset+flags, wx1, wy1, wz1} (16 bytes, written once per query by its
// own lane in tc_setup_record, read by the 8 lanes that gather the query).
//   fp32 grids  : offset = element offset of corner (x0,y0,z0) (multiple of 32), bits 0..2 = "+1 corner exists" per axis
//   fp16 x-pairs: offset = line index of (z0,y0,x0) << 2, bits 0..1 = "+1 exists" for y, z (x is folded into the line)
template <bool HALF>
__device__ __forceinline__ float4 tc_setup_record(float px, float py, float pz, int R) {
    const int S = R + 1;
    int x0, y0, z0; float wx, wy, wz; bool hx, hy, hz;
    tc_axis(px, R, x0, wx, hx);
    tc_axis(py, R, y0, wy, hy);
    tc_axis(pz, R, z0, wz, hz);
    uint32_t off;
    if constexpr (HALF) {
        // pair lines exist for x0 < R only: at the upper face (u == R) use the last pair with weight 1 on its x1
        // corner -- the same value (the reference multiplies that corner by 1 and reads nothing else)
        if (!hx) { x0 = R - 1; wx = 1.f; }
        off = ((uint32_t)((z0 * S + y0) * R + x0) << 2) | (hy ? 1u : 0u) | (hz ? 2u : 0u);
    } else {
        off = (uint32_t)(((z0 * S + y0) * S + x0) * NGLOD_F) | (hx ? 1u : 0u) | (hy ? 2u : 0u) | (hz ? 4u : 0u);
    }
    return make_float4(__uint_as_float(off), wx, wy, wz);
}

// The line loads of one (query, LOD) for channels [4c, 4c+4): 8 corner lines (fp32) or 4 x-pair lines (fp16).
template <bool HALF> struct TcLines { uint4 v[HALF ? 4 : 8]; };

template <bool HALF>
__device__ __forceinline__ void tc_issue_lines(const float* grid, int R, uint32_t pk, int c, TcLines<HALF>& t) {
    const int S = R + 1;
    if constexpr (HALF) {
        const uint4* g = reinterpret_cast<const uint4*>(grid) + ((size_t)(pk >> 2) * 8 + c);
        const int dy = (pk & 1u) ? R * 8 : 0;
        const int dz = (pk & 2u) ? S * R * 8 : 0;
        t.v[0] = __ldg(g); t.v[1] = __ldg(g + dy); t.v[2] = __ldg(g + dz); t.v[3] = __ldg(g + dz + dy);
    } else {
        const uint4* g = reinterpret_cast<const uint4*>(grid + (pk & ~31u)) + c;
        const int dx = (pk & 1u) ? NGLOD_F / 4 : 0;
        const int dy = (pk & 2u) ? S * (NGLOD_F / 4) : 0;
        const int dz = (pk & 4u) ? S * S * (NGLOD_F / 4) : 0;
        t.v[0] = __ldg(g);           t.v[1] = __ldg(g + dx);
        t.v[2] = __ldg(g + dy);      t.v[3] = __ldg(g + dy + dx);
        t.v[4] = __ldg(g + dz);      t.v[5] = __ldg(g + dz + dx);
        t.v[6] = __ldg(g + dz + dy); t.v[7] = __ldg(g + dz + dy + dx);
    }
}

__device__ __forceinline__ uint64_t u2_as_f2(uint32_t lo, uint32_t hi) { return (uint64_t)lo | ((uint64_t)hi << 32); }

// acc += trilinear sample (channels 4c..4c+3 as two packed fp32 pairs).  All arithmetic is fp32 (FFMA2 rounds each half
// exactly like a scalar fmaf); fp16 line values are converted exactly first.
template <bool HALF>
__device__ __forceinline__ void tc_consume_lines(const float4 P, const TcLines<HALF>& t, uint64_t& acc01, uint64_t& acc23) {
    const float wx1 = P.y, wy1 = P.z, wz1 = P.w;
    const float wx0 = 1.f - wx1, wy0 = 1.f - wy1, wz0 = 1.f - wz1;   // == (floor+1) - u exactly
    uint64_t s01, s23;
    if constexpr (HALF) {
        const uint64_t wx = f2_pack(wx0, wx1);
        const float wyz[4] = {wy0 * wz0, wy1 * wz0, wy0 * wz1, wy1 * wz1};
        float a, b;
        f2_unpack(f2_mul(wx, f2_pack(wyz[0], wyz[0])), a, b);
        s01 = f2_mul(h2_to_f2(t.v[0].x), f2_pack(a, a));
        s23 = f2_mul(h2_to_f2(t.v[0].y), f2_pack(a, a));
        s01 = f2_fma(h2_to_f2(t.v[0].z), f2_pack(b, b), s01);
        s23 = f2_fma(h2_to_f2(t.v[0].w), f2_pack(b, b), s23);
#pragma unroll
        for (int k = 1; k < 4; ++k) {
            f2_unpack(f2_mul(wx, f2_pack(wyz[k], wyz[k])), a, b);
            s01 = f2_fma(h2_to_f2(t.v[k].x), f2_pack(a, a), s01);
            s23 = f2_fma(h2_to_f2(t.v[k].y), f2_pack(a, a), s23);
            s01 = f2_fma(h2_to_f2(t.v[k].z), f2_pack(b, b), s01);
            s23 = f2_fma(h2_to_f2(t.v[k].w), f2_pack(b, b), s23);
        }
    } else {
        const float w00 = wx0 * wy0, w10 = wx1 * wy0, w01 = wx0 * wy1, w11 = wx1 * wy1;
        const float w[8] = {w00 * wz0, w10 * wz0, w01 * wz0, w11 * wz0, w00 * wz1, w10 * wz1, w01 * wz1, w11 * wz1};
        s01 = f2_mul(u2_as_f2(t.v[0].x, t.v[0].y), f2_pack(w[0], w[0]));
        s23 = f2_mul(u2_as_f2(t.v[0].z, t.v[0].w), f2_pack(w[0], w[0]));
#pragma unroll
        for (int k = 1; k < 8; ++k) {
            s01 = f2_fma(u2_as_f2(t.v[k].x, t.v[k].y), f2_pack(w[k], w[k]), s01);
            s23 = f2_fma(u2_as_f2(t.v[k].z, t.v[k].w), f2_pack(w[k], w[k]), s23);
        }
    }
    // running sum across LODs (OctreeSDF.py:109-110)
    acc01 = f2_add(s01, acc01);
    acc23 = f2_add(s23, acc23);
}

// Multi-LOD gather (fp32 grids): 4 queries per round, 8 lanes per corner line; the LOD count is compile-time so the
// loads of several LODs are in flight together.
template <int NL>
__device__ __forceinline__ void tc_gather_rounds(const NetDev& net, int l0, int n_live, char* a_hi, char* a_lo,
                                                 int row0, const float4* pack, const int* idx, int lane) {
    const int sub = lane >> 3, c = lane & 7;
    for (int r = 0; r * 4 < n_live; ++r) {
        const int slot = r * 4 + sub;
        if (slot < n_live) {
            const int q = idx[slot];
            const uint32_t row_off = tc_elem_offset(row0 + q, 4 * c);
            uint64_t acc01 = 0ull, acc23 = 0ull;
            if (l0 > 0) {       // more than TC_PACK_LODS grids: continue the running sum (hi + lo is exact)
                const float4 h = *reinterpret_cast<const float4*>(a_hi + row_off);
                const float4 lo = *reinterpret_cast<const float4*>(a_lo + row_off);
                acc01 = f2_pack(h.x + lo.x, h.y + lo.y); acc23 = f2_pack(h.z + lo.z, h.w + lo.w);
            }
            float4 P[NL];
            TcLines<false> t[NL];
#pragma unroll
            for (int l = 0; l < NL; ++l) {
                P[l] = pack[l * 32 + q];
                tc_issue_lines<false>(net.grids[l0 + l], net.res[l0 + l], __float_as_uint(P[l].x), c, t[l]);
            }
#pragma unroll
            for (int l = 0; l < NL; ++l) tc_consume_lines<false>(P[l], t[l], acc01, acc23);
            float4 acc;
            f2_unpack(acc01, acc.x, acc.y); f2_unpack(acc23, acc.z, acc.w);
            tc_store_split4(a_hi, a_lo, row_off, acc);
        }
    }
}

// Single-grid gather (the prefix-summed grid of the requested LOD, fp32 or fp16 x-pair lines).
// A round = queries 4r..4r+3 of the warp x 8 lanes each.  No shared-memory staging: the query's own lane keeps its
// set-up record in registers and the 8 gathering lanes fetch it with warp shuffles (a dependent LDS->LDS->LDG chain
// per round was the #2 stall after the L2 latency itself).  Software-pipelined: DEPTH rounds of line loads are in
// flight, round r+DEPTH is issued as soon as round r has been consumed.  Rounds whose 4 queries are all dead are
// skipped (only the frame tail / the last partial tile have dead lanes, so no compaction is attempted).
template <bool HALF>
__device__ __forceinline__ void tc_gather_single(const NetDev& net, const float4 rec, const unsigned live, char* a_hi,
                                                 char* a_lo, int row0, int lane) {
#ifndef NGLOD_DEPTH_HALF
#define NGLOD_DEPTH_HALF 2
#endif
#ifndef NGLOD_DEPTH_F32
#define NGLOD_DEPTH_F32 1
#endif
    constexpr int DEPTH = HALF ? NGLOD_DEPTH_HALF : NGLOD_DEPTH_F32;     // 32 data registers in flight (deeper bought nothing: profiles/README.md)
    const int sub = lane >> 3, c = lane & 7;
    const float* grid = net.grids[0];
    const int R = net.res[0];
    const uint32_t off_me = __float_as_uint(rec.x);          // dead lanes carry record 0 = corner (0,0,0), weights 0
    TcLines<HALF> t[DEPTH];
#pragma unroll
    for (int r = 0; r < DEPTH; ++r)
        if ((live >> (4 * r)) & 0xFu) tc_issue_lines<HALF>(grid, R, __shfl_sync(0xffffffffu, off_me, 4 * r + sub), c, t[r]);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        if ((live >> (4 * r)) & 0xFu) {                      // warp-uniform
            const int q = 4 * r + sub;
            float4 P;
            P.x = 0.f;
            P.y = __shfl_sync(0xffffffffu, rec.y, q);
            P.z = __shfl_sync(0xffffffffu, rec.z, q);
            P.w = __shfl_sync(0xffffffffu, rec.w, q);
            uint64_t acc01 = 0ull, acc23 = 0ull;
            tc_consume_lines<HALF>(P, t[r % DEPTH], acc01, acc23);
            if ((live >> q) & 1u) {
                float4 acc;
                f2_unpack(acc01, acc.x, acc.y); f2_unpack(acc23, acc.z, acc.w);
#ifndef NGLOD_TC_FASTSPLIT
#define NGLOD_TC_FASTSPLIT 1
#endif
#if NGLOD_TC_FASTSPLIT
                tc_store_split4_finite(a_hi, a_lo, tc_elem_offset(row0 + q, 4 * c), acc);
#else
                tc_store_split4(a_hi, a_lo, tc_elem_offset(row0 + q, 4 * c), acc);
#endif
            }
        }
        if (r + DEPTH < 8) {
            if ((live >> (4 * (r + DEPTH))) & 0xFu)
                tc_issue_lines<HALF>(grid, R, __shfl_sync(0xffffffffu, off_me, 4 * (r + DEPTH) + sub), c, t[r % DEPTH]);
        }
    }
}

// Gather for the warp's 32 queries into rows [row0, row0+32) of the group's A operand.
//   px,py,pz / active : this lane's query;   pack/idx : this warp's scratch (TC_MULTI only).
// Every lane of the warp must call (convergent).
template <int MODE>
__device__ __forceinline__ void tc_gather_rows(const NetDev& net, float px, float py, float pz, bool active,
                                               char* a_hi, char* a_lo, int row0, float4* pack, int* idx, int lane) {
    const unsigned live = __ballot_sync(0xffffffffu, active);
    const int n_live = __popc(live);
    if (n_live == 0) return;
    // K chunk 8 = {x, y, z, 1}: the query's own lane writes it (no shuffles needed later)
    if (active) tc_store_split4(a_hi, a_lo, tc_elem_offset(row0 + lane, NGLOD_F), make_float4(px, py, pz, 1.f));
    if constexpr (MODE != TC_MULTI) {
        constexpr bool HALF = MODE == TC_SINGLE_HALF;
        float4 rec = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active) rec = tc_setup_record<HALF>(px, py, pz, net.res[0]);
        tc_gather_single<HALF>(net, rec, live, a_hi, a_lo, row0, lane);
    } else {
        if (active) idx[__popc(live & ((1u << lane) - 1u))] = lane;
        for (int l0 = 0; l0 < net.num_lods; l0 += TC_PACK_LODS) {
            const int nl = min(TC_PACK_LODS, net.num_lods - l0);
            // ---- phase 1: per-LOD set-up, once per query (not once per lane of the query)
            if (active)
                for (int l = 0; l < nl; ++l) pack[l * 32 + lane] = tc_setup_record<false>(px, py, pz, net.res[l0 + l]);
            __syncwarp();
            // ---- phase 2
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
}

// d = b1 + sum_j W1[j] * relu(D[row, j]) for this lane's accumulator row (TMEM lane = 32*(warp%4) + lane).
// Four interleaved partial sums: one 128-long dependent FFMA chain is 512 cycles of pure latency per tile, which is
// what a sparsely occupied tracer tile (the frame's straggler rays) waits for every round.
__device__ __forceinline__ float tc_epilogue(uint32_t taddr_row, const float* __restrict__ w1) {
#ifndef NGLOD_TC_EPI_F2
#define NGLOD_TC_EPI_F2 1           // the two multiply-adds of a column pair as one FFMA2 (each half rounds like the scalar fmaf)
#endif
#if NGLOD_TC_EPI_F2
    uint64_t d01 = 0ull, d23 = 0ull;
#pragma unroll
    for (int cb = 0; cb < NGLOD_H / 32; ++cb) {
        float v[32];
        tmem_ld32(taddr_row + cb * 32, v);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
            const float4 w = *reinterpret_cast<const float4*>(w1 + cb * 32 + 4 * j4);
            d01 = f2_fma(f2_pack(w.x, w.y), f2_pack(fmaxf(v[4 * j4], 0.f), fmaxf(v[4 * j4 + 1], 0.f)), d01);
            d23 = f2_fma(f2_pack(w.z, w.w), f2_pack(fmaxf(v[4 * j4 + 2], 0.f), fmaxf(v[4 * j4 + 3], 0.f)), d23);
        }
    }
    float d0, d1, d2, d3;
    f2_unpack(d01, d0, d1); f2_unpack(d23, d2, d3);
    return w1[NGLOD_H] + ((d0 + d1) + (d2 + d3));
#else
    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
    for (int cb = 0; cb < NGLOD_H / 32; ++cb) {
        float v[32];
        tmem_ld32(taddr_row + cb * 32, v);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
            const float4 w = *reinterpret_cast<const float4*>(w1 + cb * 32 + 4 * j4);
            d0 = fmaf(w.x, fmaxf(v[4 * j4], 0.f), d0);
            d1 = fmaf(w.y, fmaxf(v[4 * j4 + 1], 0.f), d1);
            d2 = fmaf(w.z, fmaxf(v[4 * j4 + 2], 0.f), d2);
            d3 = fmaf(w.w, fmaxf(v[4 * j4 + 3], 0.f), d3);
        }
    }
    return w1[NGLOD_H] + ((d0 + d1) + (d2 + d3));
#endif
}

// Same sums in the same order (partial sum j mod 4 over the columns, then (d0 + d1) + (d2 + d3): bit-identical to
// tc_epilogue), arranged for a warp that shares its scheduler with gather warps: 8 columns per tcgen05.ld, the next
// chunk's load (and its W1 values) in flight while the current one is folded in, the two multiply-adds of a column pair as one FFMA2.
__device__ __forceinline__ float tc_epilogue_pipelined8(uint32_t taddr_row, const float* __restrict__ w1) {
    uint64_t d01 = 0ull, d23 = 0ull;
    uint32_t v[2][8];
    float4 w[2][2];
    tmem_ld8_async(taddr_row, v[0]);
    w[0][0] = *reinterpret_cast<const float4*>(w1);
    w[0][1] = *reinterpret_cast<const float4*>(w1 + 4);
#pragma unroll
    for (int cb = 0; cb < NGLOD_H / 8; ++cb) {
        tmem_ld_wait();
        if (cb + 1 < NGLOD_H / 8) {                      // next chunk: accumulator columns and W1 in flight during this one
            tmem_ld8_async(taddr_row + (cb + 1) * 8, v[(cb + 1) & 1]);
            w[(cb + 1) & 1][0] = *reinterpret_cast<const float4*>(w1 + (cb + 1) * 8);
            w[(cb + 1) & 1][1] = *reinterpret_cast<const float4*>(w1 + (cb + 1) * 8 + 4);
        }
        const uint32_t* u = v[cb & 1];
#pragma unroll
        for (int j4 = 0; j4 < 2; ++j4) {
            const float4 ww = w[cb & 1][j4];
            const float a0 = fmaxf(__uint_as_float(u[4 * j4]), 0.f), a1 = fmaxf(__uint_as_float(u[4 * j4 + 1]), 0.f);
            const float a2 = fmaxf(__uint_as_float(u[4 * j4 + 2]), 0.f), a3 = fmaxf(__uint_as_float(u[4 * j4 + 3]), 0.f);
            d01 = f2_fma(f2_pack(ww.x, ww.y), f2_pack(a0, a1), d01);
            d23 = f2_fma(f2_pack(ww.z, ww.w), f2_pack(a2, a3), d23);
        }
    }
    float d0, d1, d2, d3;
    f2_unpack(d01, d0, d1); f2_unpack(d23, d2, d3);
    return w1[NGLOD_H] + ((d0 + d1) + (d2 + d3));
}

__device__ __forceinline__ float tc_epilogue_pipelined16(uint32_t taddr_row, const float* __restrict__ w1) {
    uint64_t d01 = 0ull, d23 = 0ull;
    uint32_t v[2][16];
    tmem_ld16_async(taddr_row, v[0]);
#pragma unroll
    for (int cb = 0; cb < NGLOD_H / 16; ++cb) {
        tmem_ld_wait();
        if (cb + 1 < NGLOD_H / 16) tmem_ld16_async(taddr_row + (cb + 1) * 16, v[(cb + 1) & 1]);
        const uint32_t* u = v[cb & 1];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
            const float4 w = *reinterpret_cast<const float4*>(w1 + cb * 16 + 4 * j4);
            const float a0 = fmaxf(__uint_as_float(u[4 * j4]), 0.f), a1 = fmaxf(__uint_as_float(u[4 * j4 + 1]), 0.f);
            const float a2 = fmaxf(__uint_as_float(u[4 * j4 + 2]), 0.f), a3 = fmaxf(__uint_as_float(u[4 * j4 + 3]), 0.f);
            d01 = f2_fma(f2_pack(w.x, w.y), f2_pack(a0, a1), d01);
            d23 = f2_fma(f2_pack(w.z, w.w), f2_pack(a2, a3), d23);
        }
    }
    float d0, d1, d2, d3;
    f2_unpack(d01, d0, d1); f2_unpack(d23, d2, d3);
    return w1[NGLOD_H] + ((d0 + d1) + (d2 + d3));
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
    const char* b_hi; const char* b_lo;         // generic pointers to the staged B operand (W0|b0 = hi + lo exactly)
    int wq, lane, bar_id;
    uint32_t parity;
#ifdef NGLOD_TRACE_TIMING
    long long tm[8];                            // [0] refill [1] gather [2] group barrier [3] mma wait [4] epilogue [5] state
    long long t_last;
#endif
};

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

// Phase timing for experiments (profiles/exp_phases.sh): -DNGLOD_TRACE_TIMING makes the tracer accumulate clock64()
// deltas per phase into its stats buffer; compiled out otherwise.
#ifdef NGLOD_TRACE_TIMING
#define TC_TICK(i) do { const long long _t = clock64(); g.tm[i] += _t - g.t_last; g.t_last = _t; } while (0)
#else
#define TC_TICK(i) do { } while (0)
#endif

// One tile evaluation: gather -> MMA -> epilogue.  All 128 threads of the group call.
// The group barrier between "A rows written" and "MMA issued" doubles as an OR-reduction of `keep`: a persistent
// kernel whose warps must leave together passes "I still have work"; when no thread of the group has, nothing is issued
// and the function returns false (for every thread of the group alike).
template <int MODE = TC_MULTI>
__device__ __forceinline__ bool tc_group_issue(const NetDev& net, TcGroup& g, float px, float py, float pz, bool active, bool keep) {
#ifndef NGLOD_EXP_NO_GATHER      // timing experiments only (profiles/exp_parts.sh): results are garbage with these set
    tc_gather_rows<MODE>(net, px, py, pz, active, g.a_hi, g.a_lo, g.wq * 32, g.pack, g.idx, g.lane);
#endif
    fence_proxy_async_smem();                     // generic-proxy smem writes -> visible to the tensor core
    tc_fence_before_sync();                       // order the previous tile's TMEM loads before the next MMA
    TC_TICK(1);
    if (!tc_group_any(g.bar_id, keep)) return false;
    TC_TICK(2);
#ifndef NGLOD_EXP_NO_MMA
#ifndef NGLOD_TC_ISSUE_WARP
#define NGLOD_TC_ISSUE_WARP 0       // 1: the group's first warp issues as a converged warp (elect.sync); 0: lane 0 in a divergent
                                    // branch.  Measured (profiles/README.md part 4): 1 costs the 720p frame 0.90 -> 0.96 ms
#endif
#if NGLOD_TC_ISSUE_WARP
    if (g.wq == 0) {                              // warp-uniform; every warp of a group leaves the barrier converged
        tc_fence_after_sync();
        // broadcast: the addresses are the same in every lane, and the compiler keeps what derives from them uniform
        tc_issue_tile_warp(tc_warp_uniform(g.tmem_acc), tc_warp_uniform(g.a_hi_s), tc_warp_uniform(g.a_lo_s),
                           tc_warp_uniform(g.b_hi_s), tc_warp_uniform(g.b_lo_s));
        if (tc_elect_one()) tc_commit(g.mbar_s);
    }
#else
    if (g.wq == 0 && g.lane == 0) {
        tc_fence_after_sync();
        tc_issue_tile(g.tmem_acc, g.a_hi_s, g.a_lo_s, g.b_hi_s, g.b_lo_s);
        tc_commit(g.mbar_s);
    }
#endif
#endif
    return true;
}
// Second half of a tile evaluation: wait for the MMAs issued by tc_group_issue, read this lane's row, finish the decoder.
// Whatever the warp does between the two halves overlaps the tensor-core latency (~1000 cycles).
__device__ __forceinline__ float tc_group_finish(TcGroup& g, float px) {
#ifdef NGLOD_EXP_NO_MMA
    return px + g.w1[g.lane];
#endif
    mbar_wait(g.mbar_s, g.parity);
    g.parity ^= 1u;
    tc_fence_after_sync();
    TC_TICK(3);
    const float d = tc_epilogue(g.tmem_row, g.w1);
    TC_TICK(4);
    return d;
}
template <int MODE = TC_MULTI>
__device__ __forceinline__ bool tc_group_eval_any(const NetDev& net, TcGroup& g, float px, float py, float pz, bool active,
                                                  bool keep, float& d) {
    if (!tc_group_issue<MODE>(net, g, px, py, pz, active, keep)) return false;
    d = tc_group_finish(g, px);
    return true;
}

template <int MODE = TC_MULTI>
__device__ __forceinline__ float tc_group_eval(const NetDev& net, TcGroup& g, float px, float py, float pz, bool active) {
    float d = 0.f;
    tc_group_eval_any<MODE>(net, g, px, py, pz, active, true, d);
    return d;
}

__device__ __forceinline__ TcGroup tc_make_group(char* smem, int G, uint32_t tmem_base, int W = TC_WARP_SCRATCH_BYTES) {
    TcGroup g;
    const int warp = threadIdx.x >> 5;
    const int grp = warp >> 2;
    g.wq = warp & 3;
    g.lane = threadIdx.x & 31;
    g.a_hi = smem + TC_SMEM_A(grp);
    g.a_lo = g.a_hi + TC_OPERAND_BYTES;
    g.a_hi_s = smem_u32(g.a_hi);
    g.a_lo_s = smem_u32(g.a_lo);
    g.b_hi = smem + TC_SMEM_B_HI;
    g.b_lo = smem + TC_SMEM_B_LO;
    g.b_hi_s = smem_u32(g.b_hi);
    g.b_lo_s = smem_u32(g.b_lo);
    g.mbar_s = smem_u32(smem + TC_SMEM_MBAR_W(G, W) + 8 * grp);
    g.tmem_acc = tmem_base + (uint32_t)(grp * TC_N);
    g.tmem_row = g.tmem_acc + ((uint32_t)(g.wq * 32) << 16);
    char* scratch = smem + TC_SMEM_SCRATCH(G) + warp * W;
    g.pack = reinterpret_cast<float4*>(scratch);
    g.idx = reinterpret_cast<int*>(scratch + TC_PACK_LODS * 32 * 16);
    g.w1 = reinterpret_cast<const float*>(smem + TC_SMEM_W1(G));
    g.bar_id = 1 + grp;
    g.parity = 0;
    return g;
}

// Common prologue: zero the operand buffers, stage weights, init mbarriers, allocate TMEM.  Returns the TMEM base.
__device__ __forceinline__ uint32_t tc_prologue(const NetDev& net, char* smem, int G, int W = TC_WARP_SCRATCH_BYTES) {
    float wv[16];                        // one batch covers all 4737 values when blockDim >= 297 (3+ groups)
    tc_load_weights(net, 0, wv);
    for (int e = threadIdx.x; e < (TC_SMEM_SCRATCH(G) + G * 4 * W) / 16; e += blockDim.x)
        reinterpret_cast<float4*>(smem)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    tc_scatter_weights(net, smem, G, 0, wv);
    for (int base = 16 * blockDim.x; base < NGLOD_H * (NGLOD_F + 3) + 2 * NGLOD_H + 1; base += 16 * blockDim.x) {
        tc_load_weights(net, base, wv);
        tc_scatter_weights(net, smem, G, base, wv);
    }
    if (threadIdx.x == 0) {
        for (int g = 0; g < G; ++g) mbar_init(smem_u32(smem + TC_SMEM_MBAR_W(G, W) + 8 * g), 1);
        mbar_fence_init();
    }
    if (threadIdx.x < 32) tmem_alloc(smem_u32(smem + TC_SMEM_TMEMPTR_W(G, W)), 512);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    return *reinterpret_cast<volatile uint32_t*>(smem + TC_SMEM_TMEMPTR_W(G, W));
}

__device__ __forceinline__ void tc_epilogue_free(uint32_t tmem_base) {
    tc_fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem_base, 512);
}


