// nglod_b200 -- backward of OctreeSDF.sdf(x, lod) and the fused L2 training step on the tcgen05 tensor cores.
//
// Reference behaviour: autograd through sdf-net/lib/models/OctreeSDF.py:94-146 (grid_sampler_3d_backward scatter +
// Linear grads), loss of sdf-net/lib/trainer.py:317-339.  Same mathematics as sdf_backward_mma_kernel (sdf_backward.cu,
// which keeps the per-LOD and the sparse variants): the forward is recomputed, nothing is saved per query, and the only
// state a query leaves behind for the head gradients is g_d and its 128 ReLU mask bits.  Per tile of 128 queries:
//
//   GEMM1  pre[q][h]  = in[q][.] . W0ext[h][.]          kind::tf32, 3xTF32, M=128 N=128 K=40   (the forward's tile)
//   epi1   d = b1 + W1.relu(pre)  (the forward's sums in the forward's order), g_d = dL/dd,
//          MASK[q][h] = [pre > 0] as bf16 0/1 (exact) -> shared memory,   S[q][k] = g_d(q) in[q][k] as bf16 hi + lo
//   GEMM2  D2[q][t,f] = MASK[q][.] . (W1 o W0)_t[.][f]  kind::f16 (bf16), M=128 N=96 K=128: B = W1 o W0 split in three bf16 terms
//          t (24 bits) stacked along N   -> dL/dfeat[q][f] = g_d(q) (D2[q][2,f] + D2[q][1,f] + D2[q][0,f])
//   GEMM3  T[h][t,k] += MASK^T[h][.] . S_t[.][k]        kind::f16 (bf16), M=128 N=80 K=128, BOTH operands MN-major, S = hi | lo
//          stacked along N: MASK as stored for GEMM2 (K-major, K = h) IS the MN-major operand of GEMM3 (MN = h, K = q),
//          byte for byte; T lives in TMEM for the whole kernel and is read once per CTA:
//          dW0[h][k] = w1[h] T[h][k],  db0[h] = w1[h] T[h][35],  dW1[h] = sum_k W0ext[h][k] T[h][k]
//          (a chain of accumulating MMAs into one accumulator costs ~75 cycles per instruction whatever N is -- measured --
//          so the terms ride on N, 8 instructions per GEMM, and the two chains are issued interleaved)
//   epi2   dL/dfeat rows -> shared memory -> 8 lanes per corner line: 8 x red.global.add.v4.f32 per lane into ONE grid
//          (dL/d(prefix-summed grid), pushed down the LOD chain by the restriction cascade afterwards)
//
// Three gather / scatter flavours share everything else (template GM): BT_SINGLE -- the prefix-summed grid of the LOD (the
// fast path described here); BT_MULTI -- the per-LOD grids (`--no-sum-lods`, grids that do not nest): gather and scatter loop
// over the LODs; BT_SPARSE -- corner features of a sparse octree: gather and scatter walk the query's parent chain
// (sparse_core.cuh).  The two slow flavours gather synchronously (no software pipeline) from a record {x, y, z, voxel row}.
//
// Warp-specialised over a 2-stage ring (16 warps x 128 registers, one CTA per SM):
//   producers (8 warps): gather the tile's A rows exactly like the warp-specialised forward (sdf_tc.cu): records from the
//       K padding of the A rows, 8 lanes x LDG.128 per corner line, FFMA2 interpolation, hi/lo split; the last producer
//       to arrive issues GEMM1.
//   service (2 warpgroups; warpgroup g owns stage g and the tiles T = g mod 2): epi1 -> barrier -> one thread issues
//       GEMM2 + GEMM3 -> set-up of tile T+2 in the freed stage (releases the producers) -> epi2 + scatter.
// Who orders what (every hand-off is an mbarrier; `compute-sanitizer --tool racecheck` does not model mbarrier arrive /
// try_wait or tcgen05.commit and reports each of these pairs -- as it does for the forward's identical record hand-off;
// memcheck, synccheck and initcheck are clean, profiles/README.md part 4):
//   set-up writes {record, xyz} of tile T   -> producers read them              rec_full[s]  (4 service warps arrive, release)
//   producers write the A rows of tile T    -> GEMM1 reads them (async proxy)   fence.proxy.async + arrival counter, last arriver issues
//   GEMM1 done (reads of A, writes of D1)   -> epi1 reads D1, reads A_hi, overwrites A_lo with S, writes MASK     done1[s]  (tcgen05.commit)
//   epi1's MASK / S writes (all 128 threads) -> GEMM2 / GEMM3 read them          fence.proxy.async + named barrier, then one warp issues
//   GEMM2 / GEMM3 done                      -> D2 read, stage handed back (set-up of T+2 writes A rows), MASK region re-used
//                                              as the staging area                done2[s]  (tcgen05.commit)
//   staging rows: written and read by the SAME warp (__syncwarp); the next tile's epi1 rewrites the MASK only after the
//   named barrier that closes the iteration.  D1[s] / D2[s] of tile T+2 are written by MMAs issued after rec_full(T+2),
//   i.e. after every service thread's tcgen05.ld of tile T (tcgen05.fence::before_thread_sync precedes the arrive).
//
// The roof of this kernel is the reduction path: nglod_probe_scatter measures 6.2 TB/s of reduced bytes for this address
// stream (0.18 ms per 2^20 queries) whatever the launch shape -- a per-SM limit (one 512-byte RED.v4 warp instruction per
// 20-26 cycles: 50 GB/s from one CTA alone, 39 GB/s per SM with all of them busy); everything else is arranged to hide
// behind it, as far as the gather -- which shares the SM's load / store path -- lets it.
#include "sdf_tc.cuh"
#include "sdf_core.cuh"
#include "sparse_core.cuh"
#include "internal.h"
#include <cuda_bf16.h>
#include <cstdio>

namespace {

#ifndef BT_PRODUCERS
#define BT_PRODUCERS 8                       // 4, 8 or 16 gather warps
#endif
#define BT_IPT (16 / BT_PRODUCERS)             // items (8 rows = two gather rounds) per producer warp per tile
#define BT_WARPS (BT_PRODUCERS + 8)
#define BT_THREADS (BT_WARPS * 32)
#define BT_REC_COL (NGLOD_F + 4)                    // K columns 36..39 of an A row: zero in W0|b0 -> 16 free bytes per row
// shared memory (bytes)
#define BT_MASK_BYTES (128 * 128 * 2)               // MASK[q][h] bf16: (q/8)*2048 + (h/8)*128 + (q%8)*16 + (h%8)*2
#define BT_S_TERM_BYTES (5 * 2048)                  // S_t[q][k] bf16, k = 0..39: (k/8)*2048 + q*16 + (k%8)*2; S_lo follows S_hi
#define BT_STAGE_BYTES (2 * TC_OPERAND_BYTES + BT_MASK_BYTES)
#define BT_SMEM_STAGE(s) (2 * TC_OPERAND_BYTES + (s) * BT_STAGE_BYTES)            // A_hi | A_lo (later S_hi S_lo) | MASK
#define BT_SMEM_MASK(s) (BT_SMEM_STAGE(s) + 2 * TC_OPERAND_BYTES)
#define BT_SMEM_B2 (BT_SMEM_STAGE(2))               // (W1 o W0)^T [t,f][h] bf16, 3 terms t: ((32t+f)/8)*2048 + (h/8)*128 + (f%8)*16 + (h%8)*2
#define BT_B2_TERM_BYTES (4 * 2048)
#define BT_SMEM_W1 (BT_SMEM_B2 + 3 * BT_B2_TERM_BYTES)      // 128 floats + b1 (+ pad) = 528 B
#define BT_SMEM_BAR (BT_SMEM_W1 + 528)              // arrival counters[2], done1[2], done2[2], rec_full[2] (8 B each)
#define BT_SMEM_TMEMPTR (BT_SMEM_BAR + 8 * 8)
#define BT_SMEM_BYTES (BT_SMEM_TMEMPTR + 16)
static_assert(2 * BT_S_TERM_BYTES <= TC_OPERAND_BYTES, "S overwrites A_lo");
static_assert(BT_SMEM_BYTES <= 232448, "shared memory budget");
#define BT_STAGING_STRIDE 144                       // bytes per dL/dfeat row in the staging area (128 + 16: conflict-free row writes)
static_assert(128 * BT_STAGING_STRIDE <= BT_MASK_BYTES, "staging area lives in the MASK region");
// tensor memory (columns)
#define BT_TMEM_D1(s) ((uint32_t)(s) * 128u)
#define BT_TMEM_D2(s) BT_TMEM_D1(s)                 // 96 columns: D1 is dead once epi1 has read it
#define BT_TMEM_T(g) (256u + (uint32_t)(g) * 128u)  // 80 columns

// instruction descriptors, kind::f16 with bf16 operands and an fp32 accumulator, M = 128
#define BT_IDESC_BF16(N, AMN, BMN) ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(AMN) << 15) | ((uint32_t)(BMN) << 16) | \
                                    ((uint32_t)((N) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24))

// shared-memory matrix descriptor, SWIZZLE_NONE: K-major -> lbo = stride between the two 16-byte K chunks of an instruction,
// sbo = stride between 8-row groups; MN-major -> lbo = stride between groups of 8 K rows, sbo = stride between 16-byte MN chunks
__device__ __forceinline__ uint64_t bt_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;
    return d;
}
__device__ __forceinline__ void bt_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// non-blocking: has the phase with this parity completed?
__device__ __forceinline__ bool bt_mbar_test(uint32_t saddr, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(saddr), "r"(parity) : "memory");
    return ok != 0;
}

// two fp32 -> one word of two bf16 (round to nearest even), `lo` at the lower address
__device__ __forceinline__ uint32_t bt_pack_bf16(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ float bt_bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bt_bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// v (8 floats) -> bf16 hi words and bf16 words of the remainder
__device__ __forceinline__ void bt_split8(const float (&v)[8], uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        h[p] = bt_pack_bf16(v[2 * p], v[2 * p + 1]);
        l[p] = bt_pack_bf16(v[2 * p] - bt_bf16_lo(h[p]), v[2 * p + 1] - bt_bf16_hi(h[p]));
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// Phase timing for experiments (profiles/exp_bwd_timing.sh, "-DBT_TIMING"): block 0 prints the clock64() cycles one producer
// warp and one service warp of each warpgroup spent per phase; compiled out otherwise.
#ifdef BT_TIMING
#define BT_TICK(i) do { const long long _t = clock64(); tm[i] += _t - t_last; t_last = _t; } while (0)
#else
#define BT_TICK(i) do { } while (0)
#endif

enum : int { BT_SINGLE = 0, BT_MULTI = 1, BT_SPARSE = 2 };

template <bool FUSED_LOSS, int GM>
__global__ void __launch_bounds__(BT_THREADS, 1)
sdf_backward_tc_kernel(const NetDev net, const GradDev grad, const float* __restrict__ x, const long long n,
                       const float* __restrict__ grad_out, const float* __restrict__ gt, const float loss_scale,
                       float* __restrict__ loss_out, const SparseBwd sp) {
    extern __shared__ __align__(128) char smem_tc[];
    const int warp = tc_warp_uniform(threadIdx.x >> 5), lane = threadIdx.x & 31;
    float* w1s = reinterpret_cast<float*>(smem_tc + BT_SMEM_W1);
    // ---- prologue: zero the operand ring, stage W0|b0 (tf32 hi/lo), (W1 o W0)^T (3 x bf16), W1, b1; barriers; TMEM
    {
        float wv[16];
        tc_load_weights(net, 0, wv);
        for (int e = threadIdx.x; e < BT_SMEM_TMEMPTR / 16; e += blockDim.x)
            reinterpret_cast<float4*>(smem_tc)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
        tc_scatter_weights_to(net, smem_tc, w1s, 0, wv);
        for (int base = 16 * blockDim.x; base < NGLOD_H * (NGLOD_F + 3) + 2 * NGLOD_H + 1; base += 16 * blockDim.x) {
            tc_load_weights(net, base, wv);
            tc_scatter_weights_to(net, smem_tc, w1s, base, wv);
        }
        const int in_dim = net.pos_invariant ? NGLOD_F : NGLOD_F + 3;
        for (int e = threadIdx.x; e < NGLOD_H * NGLOD_F; e += blockDim.x) {
            const int h = e >> 5, f = e & 31;
            const float v = __ldg(net.w0 + h * in_dim + (net.pos_invariant ? f : f + 3)) * __ldg(net.w1 + h);
            const __nv_bfloat16 t0 = __float2bfloat16_rn(v);
            const float r1 = v - __bfloat162float(t0);
            const __nv_bfloat16 t1 = __float2bfloat16_rn(r1);
            const __nv_bfloat16 t2 = __float2bfloat16_rn(r1 - __bfloat162float(t1));
            char* dst = smem_tc + BT_SMEM_B2 + (f >> 3) * 2048 + (h >> 3) * 128 + (f & 7) * 16 + (h & 7) * 2;
            *reinterpret_cast<__nv_bfloat16*>(dst) = t0;
            *reinterpret_cast<__nv_bfloat16*>(dst + BT_B2_TERM_BYTES) = t1;
            *reinterpret_cast<__nv_bfloat16*>(dst + 2 * BT_B2_TERM_BYTES) = t2;
        }
        if (threadIdx.x == 0) {
            for (int s = 0; s < 2; ++s) {
                mbar_init(smem_u32(smem_tc + BT_SMEM_BAR + 8 * (2 + s)), 1);      // done1: GEMM1 committed
                mbar_init(smem_u32(smem_tc + BT_SMEM_BAR + 8 * (4 + s)), 1);      // done2: GEMM2 + GEMM3 committed
                mbar_init(smem_u32(smem_tc + BT_SMEM_BAR + 8 * (6 + s)), 4);      // rec_full: the stage's 4 service warps
            }
            mbar_fence_init();
        }
        if (warp == 0) tmem_alloc(smem_u32(smem_tc + BT_SMEM_TMEMPTR), 512);
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncthreads();
        tc_fence_after_sync();
    }
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_tc + BT_SMEM_TMEMPTR);
    const uint32_t bar0 = smem_u32(smem_tc + BT_SMEM_BAR);
    auto arrive_cnt = [&](int s) { return reinterpret_cast<unsigned*>(smem_tc + BT_SMEM_BAR + 8 * s); };
    auto done1_bar = [&](int s) { return bar0 + 8u * (uint32_t)(2 + s); };
    auto done2_bar = [&](int s) { return bar0 + 8u * (uint32_t)(4 + s); };
    auto rec_bar = [&](int s) { return bar0 + 8u * (uint32_t)(6 + s); };
    // CTA-local tiles: global tile gt = T * gridDim.x + blockIdx.x while gt * 128 < n; tile T lives in stage T & 1
    const long long total_tiles = (n + TC_TILE_ROWS - 1) / TC_TILE_ROWS;
    const int ntiles = (int)((total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
    const long long tile_stride = (long long)gridDim.x * TC_TILE_ROWS;
    const int R = net.res[0];
#ifdef BT_TIMING
    long long tm[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long t_last = clock64();
    const long long t_begin = t_last;
#endif

    if (warp < BT_PRODUCERS) {
        // ------------------------------------------------------------------ producers (the last to arrive issues GEMM1)
        // warp p owns rows 16p .. 16p+15 of every tile = two "items" of two gather rounds (4 queries x 8 lanes) each;
        // the 16 line loads of the next item are in flight while the current one is interpolated, split and stored
        const int sub = lane >> 3, c = lane & 7;
        const float* grid = net.grids[0];
        const uint32_t b_hi = smem_u32(smem_tc + TC_SMEM_B_HI), b_lo = smem_u32(smem_tc + TC_SMEM_B_LO);
        const int nitems = BT_IPT * ntiles;
        // this warp's rows of the tile in stage s are in place; the LAST warp to say so issues GEMM1 (whole warp in the
        // branch, one elected lane issues: the descriptors stay in uniform registers)
        auto arrive_and_issue = [&](int s) {
            fence_proxy_async_smem();
            __syncwarp();
            unsigned old = 0;
            if (lane == 0) {
                __threadfence_block();
                old = atomicAdd(arrive_cnt(s), 1u);
            }
            old = __shfl_sync(0xffffffffu, old, 0);
            if ((old & (BT_PRODUCERS - 1)) == BT_PRODUCERS - 1) {
                __threadfence_block();
                tc_fence_after_sync();
                const uint32_t a_hi_s = smem_u32(smem_tc + BT_SMEM_STAGE(s));
                tc_issue_tile_warp(tmem_base + BT_TMEM_D1(s), a_hi_s, a_hi_s + TC_OPERAND_BYTES, b_hi, b_lo);
                if (tc_elect_one()) tc_commit(done1_bar(s));
            }
            __syncwarp();
        };
        auto row_of = [&](int item) { return warp * (8 * BT_IPT) + (item % BT_IPT) * 8 + sub; };      // and row + 4
        if constexpr (GM != BT_SINGLE) {
            // per-LOD / sparse gather: record = {x, y, z, voxel row + 1 (0: inert row)}; synchronous loads
            for (int it = 0; it < nitems; ++it) {
                const int T = it / BT_IPT, s = T & 1;
                char* a_hi = smem_tc + BT_SMEM_STAGE(s);
                char* a_lo = a_hi + TC_OPERAND_BYTES;
                if (it % BT_IPT == 0) mbar_wait(rec_bar(s), (uint32_t)((T >> 1) & 1));
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int r = row_of(it) + 4 * j;
                    const float4 rec = *reinterpret_cast<const float4*>(a_hi + tc_elem_offset(r, BT_REC_COL));
                    float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
                    const int vrow1 = __float_as_int(rec.w);
                    if (vrow1 > 0) {
                        if constexpr (GM == BT_SPARSE) f = sparse_gather4(sp.sn, rec.x, rec.y, rec.z, vrow1 - 1, c);
                        else f = gather_features(net, rec.x, rec.y, rec.z, c);
                    }
                    tc_store_split4_finite(a_hi, a_lo, tc_elem_offset(r, 4 * c), f);
                }
                if (it % BT_IPT == BT_IPT - 1) arrive_and_issue(s);
            }
        }
        float4 rec0 = make_float4(0.f, 0.f, 0.f, 0.f), rec1 = rec0;
        TcLines<false> t0, t1;
        if (GM == BT_SINGLE && nitems > 0) {
            mbar_wait(rec_bar(0), 0u);
            const char* a = smem_tc + BT_SMEM_STAGE(0);
            rec0 = *reinterpret_cast<const float4*>(a + tc_elem_offset(row_of(0), BT_REC_COL));
            rec1 = *reinterpret_cast<const float4*>(a + tc_elem_offset(row_of(0) + 4, BT_REC_COL));
            tc_issue_lines<false>(grid, R, __float_as_uint(rec0.x), c, t0);
            tc_issue_lines<false>(grid, R, __float_as_uint(rec1.x), c, t1);
        }
        for (int it = 0; GM == BT_SINGLE && it < nitems; ++it) {
            const int T = it / BT_IPT, s = T & 1;
            char* a_hi = smem_tc + BT_SMEM_STAGE(s);
            char* a_lo = a_hi + TC_OPERAND_BYTES;
            const int row = row_of(it);
            const bool more = it + 1 < nitems;
            const int T1 = (it + 1) / BT_IPT, s1 = T1 & 1;
            const char* a_next = smem_tc + BT_SMEM_STAGE(s1);
            float4 nrec0 = make_float4(0.f, 0.f, 0.f, 0.f), nrec1 = nrec0;
            // the next item's loads are issued while this one is consumed -- unless it opens a tile whose set-up has not
            // arrived yet: this tile must not wait for that (its GEMM1 would, and with it the set-up it is waiting for)
            bool ahead = more;
            if (more && (it + 1) % BT_IPT == 0) {
                BT_TICK(0);
                ahead = __shfl_sync(0xffffffffu, (int)bt_mbar_test(rec_bar(s1), (uint32_t)((T1 >> 1) & 1)), 0) != 0;
            }
            if (ahead) {
                nrec0 = *reinterpret_cast<const float4*>(a_next + tc_elem_offset(row_of(it + 1), BT_REC_COL));
                nrec1 = *reinterpret_cast<const float4*>(a_next + tc_elem_offset(row_of(it + 1) + 4, BT_REC_COL));
            }
            {
                uint64_t acc01 = 0ull, acc23 = 0ull;
                tc_consume_lines<false>(rec0, t0, acc01, acc23);
                if (ahead) tc_issue_lines<false>(grid, R, __float_as_uint(nrec0.x), c, t0);
                float4 acc;
                f2_unpack(acc01, acc.x, acc.y); f2_unpack(acc23, acc.z, acc.w);
                tc_store_split4_finite(a_hi, a_lo, tc_elem_offset(row, 4 * c), acc);
            }
            {
                uint64_t acc01 = 0ull, acc23 = 0ull;
                tc_consume_lines<false>(rec1, t1, acc01, acc23);
                if (ahead) tc_issue_lines<false>(grid, R, __float_as_uint(nrec1.x), c, t1);
                float4 acc;
                f2_unpack(acc01, acc.x, acc.y); f2_unpack(acc23, acc.z, acc.w);
                tc_store_split4_finite(a_hi, a_lo, tc_elem_offset(row + 4, 4 * c), acc);
            }
            if (it % BT_IPT == BT_IPT - 1) arrive_and_issue(s);                      // this warp's rows of tile T are in place
            if (more && !ahead) {
                BT_TICK(0);
                mbar_wait(rec_bar(s1), (uint32_t)((T1 >> 1) & 1));
                BT_TICK(1);
                nrec0 = *reinterpret_cast<const float4*>(a_next + tc_elem_offset(row_of(it + 1), BT_REC_COL));
                nrec1 = *reinterpret_cast<const float4*>(a_next + tc_elem_offset(row_of(it + 1) + 4, BT_REC_COL));
                tc_issue_lines<false>(grid, R, __float_as_uint(nrec0.x), c, t0);
                tc_issue_lines<false>(grid, R, __float_as_uint(nrec1.x), c, t1);
            }
            rec0 = nrec0; rec1 = nrec1;
        }
    } else {
        // ------------------------------------------------------------------ service warpgroup g: stage g, tiles T = g (mod 2)
        const int g = (warp - BT_PRODUCERS) >> 2;
        const int ew = warp & 3;                          // TMEM lane quarter = rows 32 ew .. 32 ew + 31
        const int row = ew * 32 + lane;
        char* a_hi = smem_tc + BT_SMEM_STAGE(g);
        char* a_lo = a_hi + TC_OPERAND_BYTES;
        char* s_hi = a_lo;                                // S overwrites A_lo (dead once GEMM1 has completed)
        char* s_lo = a_lo + BT_S_TERM_BYTES;
        char* mask = smem_tc + BT_SMEM_MASK(g);
        char* staging = mask;                             // dL/dfeat rows (the MASK is dead once GEMM2 / GEMM3 have completed)
        const uint32_t rec_off = tc_elem_offset(row, BT_REC_COL), xyz_off = tc_elem_offset(row, NGLOD_F);
        const uint32_t lane_sel = (uint32_t)(ew * 32) << 16;
        const uint32_t t_d1 = tmem_base + BT_TMEM_D1(g) + lane_sel, t_d2 = tmem_base + BT_TMEM_D2(g) + lane_sel;
        const long long i_first = (long long)blockIdx.x * TC_TILE_ROWS + row;
        // small grids: private copies per CTA group (the L2 atomic units serialise per address), folded by the launcher
        float* ggrid = grad.priv ? grad.priv + (size_t)(blockIdx.x % grad.priv_copies) * grad.priv_stride : grad.grids[0];
        float acc_b1 = 0.f, acc_loss = 0.f;
        int nv = 0;                                       // BT_SPARSE: voxel row of the prefetched query, -1 outside the octree
        auto load_xyz = [&](int T, float& px, float& py, float& pz) {
            const long long i = i_first + (long long)T * tile_stride;
            px = py = pz = 0.f;
            if (T < ntiles && i < n) {
                px = __ldg(x + 3 * i); py = __ldg(x + 3 * i + 1); pz = __ldg(x + 3 * i + 2);
                if constexpr (GM == BT_SPARSE) nv = __ldg(sp.pidx + i);
            }
        };
        auto setup = [&](int T, float px, float py, float pz) -> float4 {
            const long long i = i_first + (long long)T * tile_stride;
            float4 rec = make_float4(0.f, 0.f, 0.f, 0.f);
            if constexpr (GM == BT_SINGLE) {
                // rows past n carry record 0 (corner 0, weights 0): harmless loads; their g_d is 0, so they add nothing
                if (i < n) rec = tc_setup_record<false>(px, py, pz, R);
            } else {
                // {x, y, z, voxel row + 1}; 0 = inert row (past n, or a point outside the octree): no gather, no loss, no gradient.
                // The word is a small integer, i.e. a finite (denormal) float: the MMA multiplies it by W0's zero padding
                int vrow1 = 0;
                if (i < n) vrow1 = GM == BT_SPARSE ? (nv >= 0 ? sp.sn.vox_off + nv + 1 : 0) : 1;
                rec = make_float4(px, py, pz, __int_as_float(vrow1));
            }
            *reinterpret_cast<float4*>(a_hi + rec_off) = rec;
            tc_store_split4(a_hi, a_lo, xyz_off, make_float4(px, py, pz, 1.f));
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(rec_bar(g));
            return rec;
        };
        float nx, ny, nz;
        float4 rec_cur = make_float4(0.f, 0.f, 0.f, 0.f);
        load_xyz(g, nx, ny, nz);
        if (g < ntiles) rec_cur = setup(g, nx, ny, nz);
        for (int T = g; T < ntiles; T += 2) {
            const uint32_t par = (uint32_t)((T >> 1) & 1);
            const long long i = i_first + (long long)T * tile_stride;
            const bool active = i < n && (GM == BT_SINGLE || __float_as_int(rec_cur.w) > 0);
            float up = 0.f;                                // the label (fused loss) or the upstream gradient
            if (active) up = __ldg((FUSED_LOSS ? gt : grad_out) + i);
            load_xyz(T + 2, nx, ny, nz);
            BT_TICK(0);
            mbar_wait(done1_bar(g), par);
            tc_fence_after_sync();
            BT_TICK(1);
            // ---- epi1: d (the forward kernel's sums in the forward kernel's order), ReLU mask -> MASK rows
            float d;
            {
                uint64_t d01 = 0ull, d23 = 0ull;
                uint32_t v[2][16];
                tmem_ld16_async(t_d1, v[0]);
                char* mrow = mask + (row >> 3) * 2048 + (row & 7) * 16;
#pragma unroll
                for (int cb = 0; cb < NGLOD_H / 16; ++cb) {
                    tmem_ld_wait();
                    if (cb + 1 < NGLOD_H / 16) tmem_ld16_async(t_d1 + (cb + 1) * 16, v[(cb + 1) & 1]);
                    const uint32_t* u = v[cb & 1];
                    uint32_t mw[8];
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        const float4 w = *reinterpret_cast<const float4*>(w1s + cb * 16 + 4 * j4);
                        const float p0 = __uint_as_float(u[4 * j4]), p1 = __uint_as_float(u[4 * j4 + 1]);
                        const float p2 = __uint_as_float(u[4 * j4 + 2]), p3 = __uint_as_float(u[4 * j4 + 3]);
                        d01 = f2_fma(f2_pack(w.x, w.y), f2_pack(fmaxf(p0, 0.f), fmaxf(p1, 0.f)), d01);
                        d23 = f2_fma(f2_pack(w.z, w.w), f2_pack(fmaxf(p2, 0.f), fmaxf(p3, 0.f)), d23);
                        mw[2 * j4] = (p0 > 0.f ? 0x3F80u : 0u) | (p1 > 0.f ? 0x3F800000u : 0u);
                        mw[2 * j4 + 1] = (p2 > 0.f ? 0x3F80u : 0u) | (p3 > 0.f ? 0x3F800000u : 0u);
                    }
                    *reinterpret_cast<uint4*>(mrow + (2 * cb) * 128) = make_uint4(mw[0], mw[1], mw[2], mw[3]);
                    *reinterpret_cast<uint4*>(mrow + (2 * cb + 1) * 128) = make_uint4(mw[4], mw[5], mw[6], mw[7]);
                }
                float d0, d1, d2, d3;
                f2_unpack(d01, d0, d1); f2_unpack(d23, d2, d3);
                d = w1s[NGLOD_H] + ((d0 + d1) + (d2 + d3));
            }
            BT_TICK(2);
            float gd = 0.f;
            if (active) {
                if (FUSED_LOSS) {
                    const float diff = d - up;
                    acc_loss = fmaf(diff * diff, loss_scale, acc_loss);
                    gd = 2.f * diff * loss_scale;
                } else {
                    gd = up;
                }
            }
            acc_b1 += gd;
            // ---- S[q][k] = g_d in[q][k] (in as GEMM1 saw its leading term: the TF32 high part), bf16 hi + lo
#pragma unroll
            for (int c5 = 0; c5 < 5; ++c5) {
                const float4 a = *reinterpret_cast<const float4*>(a_hi + tc_elem_offset(row, 8 * c5));
                float4 b = make_float4(0.f, 0.f, 0.f, 0.f);             // columns 36..39 hold the gather record, not data
                if (c5 < 4) b = *reinterpret_cast<const float4*>(a_hi + tc_elem_offset(row, 8 * c5 + 4));
                const float v[8] = {gd * a.x, gd * a.y, gd * a.z, gd * a.w, gd * b.x, gd * b.y, gd * b.z, gd * b.w};
                uint4 hi, lo;
                bt_split8(v, hi, lo);
                *reinterpret_cast<uint4*>(s_hi + c5 * 2048 + row * 16) = hi;
                *reinterpret_cast<uint4*>(s_lo + c5 * 2048 + row * 16) = lo;
            }
            fence_proxy_async_smem();
            tc_fence_before_sync();
            BT_TICK(3);
            named_bar_sync(1 + g, 128);
            BT_TICK(4);
            if (ew == 0) {                                 // the whole warp; one elected lane issues
                tc_fence_after_sync();
                const uint32_t m_s = smem_u32(mask), b2_s = smem_u32(smem_tc + BT_SMEM_B2);
                const uint32_t s_hi_s = smem_u32(s_hi);
                // GEMM2: D2[q][t,f] = MASK (K-major: 8-row groups 2048 B apart, K chunks 128 B apart) x B2 (K-major, same strides)
                // GEMM3: T[h][t,k] += MASK^T (MN-major: h chunks 128 B apart, groups of 8 queries 2048 B apart)
                //                     x S (MN-major: k chunks 2048 B apart, groups of 8 queries 128 B apart)
                // two independent accumulation chains, issued alternately
#pragma unroll
                for (int ks = 0; ks < NGLOD_H / 16; ++ks) {
                    const uint64_t a2 = bt_desc(m_s + ks * 256, 128, 2048), b2 = bt_desc(b2_s + ks * 256, 128, 2048);
                    const uint64_t a3 = bt_desc(m_s + ks * 4096, 2048, 128), b3 = bt_desc(s_hi_s + ks * 256, 128, 2048);
                    if (tc_elect_one()) {
                        bt_mma_bf16(tmem_base + BT_TMEM_D2(g), a2, b2, BT_IDESC_BF16(96, 0, 0), ks != 0);
                        bt_mma_bf16(tmem_base + BT_TMEM_T(g), a3, b3, BT_IDESC_BF16(80, 1, 1), (T != g) | (ks != 0));
                    }
                }
                if (tc_elect_one()) tc_commit(done2_bar(g));
            }
            BT_TICK(5);
            mbar_wait(done2_bar(g), par);
            tc_fence_after_sync();
            BT_TICK(6);
            // ---- epi2: dL/dfeat[q][.] = g_d (D2 term 2 + term 1 + term 0), read before the stage is handed back (D2 lives in
            //      D1's columns, which GEMM1 of tile T+2 overwrites)
            float gin[32];
            {
                float v[32];
                tmem_ld32(t_d2 + 64, gin);
                tmem_ld32(t_d2 + 32, v);
#pragma unroll
                for (int k = 0; k < 32; ++k) gin[k] += v[k];
                tmem_ld32(t_d2, v);
#pragma unroll
                for (int k = 0; k < 32; ++k) gin[k] = gd * (gin[k] + v[k]);
                tc_fence_before_sync();
            }
            // ---- the stage's operands are dead: set up tile T+2 in it (the producers fill it while this warpgroup scatters)
            float4 rec_next = make_float4(0.f, 0.f, 0.f, 0.f);
            if (T + 2 < ntiles) rec_next = setup(T + 2, nx, ny, nz);
            {
                float4* dst = reinterpret_cast<float4*>(staging + row * BT_STAGING_STRIDE);
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4) dst[k4] = make_float4(gin[4 * k4], gin[4 * k4 + 1], gin[4 * k4 + 2], gin[4 * k4 + 3]);
            }
            __syncwarp();
            BT_TICK(7);
            // ---- scatter: this warp's 32 rows, 4 queries per round x 8 lanes per corner line
#ifndef BT_EXP_NOSCATTER
            if (GM == BT_SPARSE ? sp.grad_cf != nullptr : (GM == BT_MULTI || ggrid != nullptr)) {
                const unsigned live = __ballot_sync(0xffffffffu, active && gd != 0.f);
                const int sub = lane >> 3, c = lane & 7;
                const int S = R + 1;
#pragma unroll 2
                for (int r = 0; r < 8; ++r) {
                    if (!((live >> (4 * r)) & 0xFu)) continue;             // warp-uniform
                    const int q = 4 * r + sub;
                    const float rx = __shfl_sync(0xffffffffu, rec_cur.x, q);
                    const float ry = __shfl_sync(0xffffffffu, rec_cur.y, q);
                    const float rz = __shfl_sync(0xffffffffu, rec_cur.z, q);
                    const float rw = __shfl_sync(0xffffffffu, rec_cur.w, q);
                    if ((live >> q) & 1u) {
                        const float4 gq = *reinterpret_cast<const float4*>(staging + (ew * 32 + q) * BT_STAGING_STRIDE + 16 * c);
                        if constexpr (GM == BT_SINGLE) {
                            const uint32_t pk = __float_as_uint(rx);
                            const float wx1 = ry, wy1 = rz, wz1 = rw;
                            const float wx0 = 1.f - wx1, wy0 = 1.f - wy1, wz0 = 1.f - wz1;
                            const float w00 = wx0 * wy0, w10 = wx1 * wy0, w01 = wx0 * wy1, w11 = wx1 * wy1;
                            const float w[8] = {w00 * wz0, w10 * wz0, w01 * wz0, w11 * wz0, w00 * wz1, w10 * wz1, w01 * wz1, w11 * wz1};
                            float* base = ggrid + (pk & ~31u) + 4 * c;
                            const int dx = (pk & 1u) ? NGLOD_F : 0;
                            const int dy = (pk & 2u) ? S * NGLOD_F : 0;
                            const int dz = (pk & 4u) ? S * S * NGLOD_F : 0;
                            const int offs[8] = {0, dx, dy, dy + dx, dz, dz + dx, dz + dy, dz + dy + dx};
#pragma unroll
                            for (int k = 0; k < 8; ++k)
                                red_add_v4(base + offs[k], gq.x * w[k], gq.y * w[k], gq.z * w[k], gq.w * w[k]);
                        } else if constexpr (GM == BT_SPARSE) {
                            sparse_scatter4(sp.sn, sp.grad_cf, rx, ry, rz, __float_as_int(rw) - 1, c, gq);
                        } else {
#pragma unroll
                            for (int l = 0; l < NGLOD_MAX_LODS; ++l) {
                                if (l >= net.num_lods) break;
                                float* gg = grad.grids[l];
                                if (!gg) continue;
                                LodSetup ls;
                                lod_setup(rx, ry, rz, net.res[l], ls);
#pragma unroll
                                for (int k = 0; k < 8; ++k)
                                    red_add_v4(gg + ls.off[k] + 4 * c, gq.x * ls.w[k], gq.y * ls.w[k], gq.z * ls.w[k], gq.w * ls.w[k]);
                            }
                        }
                    }
                }
            }
#endif
            rec_cur = rec_next;
            BT_TICK(8);
            named_bar_sync(1 + g, 128);        // every warp is done with the staging rows before epi1 of T+2 rewrites the MASK
        }
        BT_TICK(9);
        // ---- flush the head gradients of this warpgroup's tiles: thread = hidden unit
        if (g < ntiles) {
            const int h = row;
            const uint32_t t_T = tmem_base + BT_TMEM_T(g) + lane_sel;
            float tv[40];
            {
                float v[32];
                uint32_t u[2][8], w[16];
                tmem_ld32(t_T, v);                       // hi part, k = 0..31
                tmem_ld8_async(t_T + 32, u[0]);          // hi part, k = 32..39
                tmem_ld8_async(t_T + 40, u[1]);          // lo part, k = 0..7
                tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 32; ++k) tv[k] = v[k];
#pragma unroll
                for (int k = 0; k < 8; ++k) { tv[32 + k] = __uint_as_float(u[0][k]); tv[k] += __uint_as_float(u[1][k]); }
                tmem_ld16_async(t_T + 48, w);            // lo part, k = 8..23
                tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 16; ++k) tv[8 + k] += __uint_as_float(w[k]);
                tmem_ld16_async(t_T + 64, w);            // lo part, k = 24..39
                tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 16; ++k) tv[24 + k] += __uint_as_float(w[k]);
            }
            tc_fence_before_sync();
            const int in_dim = net.pos_invariant ? NGLOD_F : NGLOD_F + 3;
            const float w1h = w1s[h];
            float dw1 = 0.f;
#pragma unroll
            for (int k = 0; k < NGLOD_KPAD; ++k) {
                const uint32_t off = tc_elem_offset(h, k);
                const float w0e = *reinterpret_cast<const float*>(smem_tc + TC_SMEM_B_HI + off) +
                                  *reinterpret_cast<const float*>(smem_tc + TC_SMEM_B_LO + off);       // W0ext[h][k], exactly
                dw1 = fmaf(w0e, tv[k], dw1);
                const float v = w1h * tv[k];
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
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            acc_b1 += __shfl_xor_sync(0xffffffffu, acc_b1, o);
            acc_loss += __shfl_xor_sync(0xffffffffu, acc_loss, o);
        }
        if (lane == 0 && g < ntiles) {
            if (grad.b1) atomicAdd(grad.b1, acc_b1);
            if (FUSED_LOSS && loss_out) atomicAdd(loss_out, acc_loss);
        }
    }
#ifdef BT_TIMING
    if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == BT_PRODUCERS || warp == BT_PRODUCERS + 4)) {
        const long long tot = clock64() - t_begin;
        if (warp == 0)
            printf("producer: tiles %d total %lld | work %lld wait_rec %lld\n", ntiles, tot, tm[0], tm[1]);
        else
            printf("service %d: total %lld | other %lld wait_done1 %lld epi1 %lld S %lld bar %lld issue %lld wait_done2 %lld setup+epi2 %lld "
                   "scatter %lld endbar %lld\n", (warp - BT_PRODUCERS) >> 2, tot, tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], tm[6], tm[7], tm[8], tm[9]);
    }
#endif
    tc_epilogue_free(tmem_base);
}

template <bool FUSED_LOSS, int GM>
int launch_bt(const NetDev& nd, const GradDev& gd, const float* x, long long n, const float* grad_out, const float* gt,
              float loss_scale, float* loss_out, const SparseBwd& sp, cudaStream_t st) {
    long long grid = nglod_sm_count();
    const long long want = (n + TC_TILE_ROWS - 1) / TC_TILE_ROWS;
    if (want < grid) grid = want;
    auto kern = sdf_backward_tc_kernel<FUSED_LOSS, GM>;
    NGLOD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, BT_SMEM_BYTES));
    kern<<<(int)grid, BT_THREADS, BT_SMEM_BYTES, st>>>(nd, gd, x, n, grad_out, gt, loss_scale, loss_out, sp);
    return (int)cudaGetLastError();
}

}  // namespace

int nglod_launch_sdf_backward_tc(const NetDev& nd, const GradDev& gd, const float* x, long long n, const float* grad_out,
                                 const float* gt, float loss_scale, float* loss_out, bool fused_loss, cudaStream_t st,
                                 const SparseBwd* sp) {
    const SparseBwd none{};
    if (sp) {
        return fused_loss ? launch_bt<true, BT_SPARSE>(sp->sn.dec, gd, x, n, grad_out, gt, loss_scale, loss_out, *sp, st)
                          : launch_bt<false, BT_SPARSE>(sp->sn.dec, gd, x, n, grad_out, gt, loss_scale, loss_out, *sp, st);
    }
    if (nd.num_lods == 1) {         // one grid to gather from and scatter into: the prefix-summed grid, or a one-level net
        return fused_loss ? launch_bt<true, BT_SINGLE>(nd, gd, x, n, grad_out, gt, loss_scale, loss_out, none, st)
                          : launch_bt<false, BT_SINGLE>(nd, gd, x, n, grad_out, gt, loss_scale, loss_out, none, st);
    }
    return fused_loss ? launch_bt<true, BT_MULTI>(nd, gd, x, n, grad_out, gt, loss_scale, loss_out, none, st)
                      : launch_bt<false, BT_MULTI>(nd, gd, x, n, grad_out, gt, loss_scale, loss_out, none, st);
}
