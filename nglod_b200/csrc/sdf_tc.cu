// nglod_b200 -- OctreeSDF.sdf forward with the decoder on tcgen05 tensor cores (3xTF32, TMEM accumulator).
// See sdf_tc.cuh / tc_common.cuh for the scheme.  Reference behaviour: sdf-net/lib/models/OctreeSDF.py:94-146.
#include "sdf_tc.cuh"
#include "internal.h"
#include <cstdio>

namespace {

#ifndef NGLOD_FWD_GROUPS
#define NGLOD_FWD_GROUPS 3          // multi-LOD gather: groups of 128 threads per CTA, one CTA per SM
#endif
#ifndef NGLOD_FWD_GROUPS_SINGLE
#define NGLOD_FWD_GROUPS_SINGLE 4   // single-grid gather: no scratch, <=128 registers -> 16 warps, all 512 TMEM columns
#endif
constexpr int fwd_groups(int mode) { return mode == TC_MULTI ? NGLOD_FWD_GROUPS : NGLOD_FWD_GROUPS_SINGLE; }
constexpr int fwd_smem(int mode) { return TC_SMEM_BYTES_W(fwd_groups(mode), tc_mode_scratch(mode)); }

template <int MODE>
__global__ void __launch_bounds__(fwd_groups(MODE) * TCG_THREADS, 1)
sdf_forward_tc_kernel(const NetDev net, const float* __restrict__ x, const long long n, float* __restrict__ out) {
    extern __shared__ __align__(128) char smem_tc[];
    constexpr int G = fwd_groups(MODE), W = tc_mode_scratch(MODE);
    if constexpr (MODE != TC_MULTI) {
        // A big batch touches every line of the grid many times, but after a cold start (new weights / flushed L2) the
        // first tiles would fetch their lines from DRAM one dependent round at a time (~10 us before the L2 is warm):
        // stream the grid into L2 at HBM speed while the prologue stages the weights.
        if (n >= (1ll << 17)) {
            const long long S = net.res[0] + 1;
            const long long bytes = MODE == TC_SINGLE_HALF ? S * S * net.res[0] * 128 : S * S * S * NGLOD_F * 4;
            const char* base = reinterpret_cast<const char*>(net.grids[0]);
            for (long long off = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 128; off < bytes;
                 off += (long long)gridDim.x * blockDim.x * 128)
                asm volatile("prefetch.global.L2 [%0];" :: "l"(base + off));
        }
    }
    const uint32_t tmem_base = tc_prologue(net, smem_tc, G, W);
    TcGroup g = tc_make_group(smem_tc, G, tmem_base, W);
    const long long ggroup = (long long)blockIdx.x * G + (threadIdx.x >> 7);
    const long long ngroups = (long long)gridDim.x * G;
    // the next tile's coordinates are fetched (from DRAM) while the current tile is evaluated
    long long tile = ggroup;
    long long i = tile * TCG_THREADS + g.wq * 32 + g.lane;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (i < n) { px = __ldg(x + 3 * i); py = __ldg(x + 3 * i + 1); pz = __ldg(x + 3 * i + 2); }
    for (; tile * TCG_THREADS < n; tile += ngroups) {
        const long long i_next = i + ngroups * TCG_THREADS;
        float nx = 0.f, ny = 0.f, nz = 0.f;
        if (i_next < n) { nx = __ldg(x + 3 * i_next); ny = __ldg(x + 3 * i_next + 1); nz = __ldg(x + 3 * i_next + 2); }
        const bool active = i < n;
        const float d = tc_group_eval<MODE>(net, g, px, py, pz, active);
        if (active) out[i] = d;
        i = i_next; px = nx; py = ny; pz = nz;
    }
    tc_epilogue_free(tmem_base);
}

template <int MODE>
int launch_fwd(const NetDev& nd, const float* x, long long n, float* out, cudaStream_t st) {
    auto kern = sdf_forward_tc_kernel<MODE>;
    constexpr int threads = fwd_groups(MODE) * TCG_THREADS, smem = fwd_smem(MODE);
    static_assert(smem <= 232448, "shared memory budget");
    NGLOD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    long long grid = nglod_sm_count();
    const long long want = (n + threads - 1) / threads;
    if (want < grid) grid = want;
    kern<<<(int)grid, threads, smem, st>>>(nd, x, n, out);
    return (int)cudaGetLastError();
}


// ------------------------------------------------------------------------------------------------------------------
// Warp-specialised forward (single-grid gathers).  The kernel above runs set-up -> gather -> barrier -> 15 UMMAs ->
// epilogue serially inside each 4-warp group, so a warp has loads in flight only ~40 % of the time and the SM pulls
// ~35 B/clk out of L2 where the probe (probe.cu) measures ~69 B/clk for the same address stream.  Here the phases run
// on different warps and overlap tile by tile over a 4-stage ring of A operands / TMEM accumulators:
//   service (2 warpgroups of 4 warps, the HIGHEST warp ids: the issue arbiter prefers them and their work is short and
//       latency-critical; warpgroup g serves the tiles T = g (mod 2)):
//       epilogue of tile T: waits mma_done[s], tcgen05.ld its 32 rows, d = b1 + W1.relu(.), stores out.  One warpgroup
//         needs ~2 800 cycles per tile next to the gather warps (measured) -- two of them keep up with the producers;
//       set-up of tile T+4 (same stage, just freed): one query per lane (coalesced coordinate loads, prefetched) -> the
//         16-byte gather record {corner offset | flags, wx, wy, wz} and the {x, y, z, 1} K-chunk, written straight into
//         the tile's A rows: the record lives in the K-padding chunk (columns 36..39), which the MMA multiplies by the
//         zero padding of W0 -- finite x 0, contributes nothing -- so it costs no shared memory -> mbarrier rec_full[s]
//   producers (16 warps): ALL of them work on the same tile: warp p owns rows 8p..8p+7 = two gather rounds of 4 queries
//       x 8 lanes: record by one LDS.128, 8 corner lines per query as 8 lanes x LDG.128, FFMA2 interpolation, hi/lo split,
//       A rows.  A tile is filled in two load latencies, so a stage is re-used only four tile times after its MMAs.
//   MMA issue: whichever producer warp arrives LAST at the tile (monotonic shared counter, 16 arrivals per tile): its
//       lane 0 issues the 15 tcgen05.mma of the tile into TMEM accumulator s and tcgen05.commit -> mbarrier mma_done[s].
//       No warp is parked on a barrier for it and the MMAs start the moment the tile is full.
// Ordering that needs no barrier of its own: the set-up of tile T+4 (same stage, same accumulator as T) follows the
// epilogue of tile T in the same warps, and the producers fill T+4 only after that set-up.
// Registers follow the roles (setmaxnreg moves registers inside the CTA's own allocation: 24 warps x 80 at launch):
// the two service warpgroups drop to 48, the four producer warpgroups grow to 96 (64 of them hold line loads in flight):
// 256 x 48 + 512 x 96 = 768 x 80 exactly.
// Tiles are assigned statically (CTA-local tile T = global tile T*gridDim.x + blockIdx.x): every role derives the same
// mapping, nothing is communicated but the barriers.  Same arithmetic as the kernel above (tc_setup_record /
// tc_issue_lines / tc_consume_lines / tc_issue_tile / tc_epilogue): results are bit-identical.
#define WS_PRODUCERS 16
#define WS_STAGES 4
#define WS_WARPS (WS_PRODUCERS + 8)
#define WS_REGS_PRODUCER 96
#define WS_REGS_SERVICE 48
#define WS_THREADS (WS_WARPS * 32)
#define WS_REC_COL (NGLOD_F + 4)                      // K columns 36..39: zero in W0|b0 -> free 16 bytes per A row
#define WS_SMEM_A(s) (2 * TC_OPERAND_BYTES + (s) * 2 * TC_OPERAND_BYTES)
#define WS_SMEM_W1 TC_SMEM_W1(WS_STAGES)
#define WS_SMEM_BAR (WS_SMEM_W1 + 528)                 // arrival counters[4] (u32, padded to 8 B), mma_done[4], rec_full[4]
#define WS_SMEM_TMEMPTR (WS_SMEM_BAR + 3 * WS_STAGES * 8)
#define WS_SMEM_BYTES (WS_SMEM_TMEMPTR + 16)
static_assert(WS_SMEM_BYTES <= 232448, "shared memory budget");
static_assert(256 * WS_REGS_SERVICE + 512 * WS_REGS_PRODUCER <= WS_THREADS * 80,
              "setmaxnreg.inc spins forever when the CTA's register pool cannot cover it");

template <int MODE>
__global__ void __launch_bounds__(WS_THREADS, 1)
sdf_forward_ws_kernel(const NetDev net, const float* __restrict__ x, const long long n, float* __restrict__ out) {
    extern __shared__ __align__(128) char smem_tc[];
    constexpr bool HALF = MODE == TC_SINGLE_HALF;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (n >= (1ll << 17)) {        // cold start: stream the grid into L2 at HBM speed while the prologue runs
        const long long S = net.res[0] + 1;
        const long long bytes = HALF ? S * S * net.res[0] * 128 : S * S * S * NGLOD_F * 4;
        const char* base = reinterpret_cast<const char*>(net.grids[0]);
        for (long long off = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 128; off < bytes;
             off += (long long)gridDim.x * blockDim.x * 128)
            asm volatile("prefetch.global.L2 [%0];" :: "l"(base + off));
    }
    // ---- prologue: zero the operand ring, stage W0|b0 (hi/lo), W1, b1; barriers; TMEM
    {
        float wv[16];
        tc_load_weights(net, 0, wv);
        for (int e = threadIdx.x; e < WS_SMEM_BAR / 16; e += blockDim.x)
            reinterpret_cast<float4*>(smem_tc)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
        tc_scatter_weights(net, smem_tc, WS_STAGES, 0, wv);
        for (int base = 16 * blockDim.x; base < NGLOD_H * (NGLOD_F + 3) + 2 * NGLOD_H + 1; base += 16 * blockDim.x) {
            tc_load_weights(net, base, wv);
            tc_scatter_weights(net, smem_tc, WS_STAGES, base, wv);
        }
        if (threadIdx.x == 0) {
            for (int s = 0; s < WS_STAGES; ++s) {
                *reinterpret_cast<volatile unsigned long long*>(smem_tc + WS_SMEM_BAR + 8 * s) = 0ull;    // arrival counter
                mbar_init(smem_u32(smem_tc + WS_SMEM_BAR + 8 * (WS_STAGES + s)), 1);         // mma_done: tcgen05.commit
                mbar_init(smem_u32(smem_tc + WS_SMEM_BAR + 8 * (2 * WS_STAGES + s)), 4);     // rec_full: the tile's 4 service warps
            }
            mbar_fence_init();
        }
        if (warp == 0) tmem_alloc(smem_u32(smem_tc + WS_SMEM_TMEMPTR), 512);
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncthreads();
        tc_fence_after_sync();
    }
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_tc + WS_SMEM_TMEMPTR);
    const uint32_t bar0 = smem_u32(smem_tc + WS_SMEM_BAR);
    auto arrive_cnt = [&](int s) { return reinterpret_cast<unsigned*>(smem_tc + WS_SMEM_BAR + 8 * s); };
    auto done_bar = [&](int s) { return bar0 + 8u * (uint32_t)(WS_STAGES + s); };
    auto rec_bar = [&](int s) { return bar0 + 8u * (uint32_t)(2 * WS_STAGES + s); };
    // CTA-local tiles: global tile gt = T * gridDim.x + blockIdx.x while gt * 128 < n
    const long long total_tiles = (n + TC_TILE_ROWS - 1) / TC_TILE_ROWS;
    const int ntiles = (int)((total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
    const long long tile_stride = (long long)gridDim.x * TC_TILE_ROWS;

    if (warp < WS_PRODUCERS) {
        // ------------------------------------------------------------------ producers (the last to arrive issues the MMAs)
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(WS_REGS_PRODUCER));
        const int sub = lane >> 3, c = lane & 7;
        const float* grid = net.grids[0];
        const int R = net.res[0];
        const uint32_t b_hi = smem_u32(smem_tc + TC_SMEM_B_HI), b_lo = smem_u32(smem_tc + TC_SMEM_B_LO);
        const int row0 = warp * 8 + sub;                  // this lane's two rows of every tile: row0, row0 + 4
        const uint32_t rec_off0 = tc_elem_offset(row0, WS_REC_COL), rec_off1 = tc_elem_offset(row0 + 4, WS_REC_COL);
        const uint32_t st_off0 = tc_elem_offset(row0, 4 * c), st_off1 = tc_elem_offset(row0 + 4, 4 * c);
        // software-pipelined across tiles: the 16 line loads of tile T+1 are issued while tile T is interpolated and
        // stored, so the L2 latency hides behind ~200 instructions of arithmetic instead of being waited for
        float4 rec0 = make_float4(0.f, 0.f, 0.f, 0.f), rec1 = rec0;
        TcLines<HALF> t0, t1;
        if (ntiles > 0) {
            mbar_wait(rec_bar(0), 0u);
            rec0 = *reinterpret_cast<const float4*>(smem_tc + WS_SMEM_A(0) + rec_off0);
            rec1 = *reinterpret_cast<const float4*>(smem_tc + WS_SMEM_A(0) + rec_off1);
            tc_issue_lines<HALF>(grid, R, __float_as_uint(rec0.x), c, t0);
            tc_issue_lines<HALF>(grid, R, __float_as_uint(rec1.x), c, t1);
        }
        for (int T = 0; T < ntiles; ++T) {
            const int s = T & (WS_STAGES - 1);
            char* a_hi = smem_tc + WS_SMEM_A(s);
            char* a_lo = a_hi + TC_OPERAND_BYTES;
            const bool more = T + 1 < ntiles;
            float4 nrec0 = make_float4(0.f, 0.f, 0.f, 0.f), nrec1 = nrec0;
            if (more) {                                   // records of the next tile (set up three tiles ahead of the epilogue)
                const int s1 = (T + 1) & (WS_STAGES - 1), k1 = (T + 1) / WS_STAGES;
                mbar_wait(rec_bar(s1), (uint32_t)(k1 & 1));
                nrec0 = *reinterpret_cast<const float4*>(smem_tc + WS_SMEM_A(s1) + rec_off0);
                nrec1 = *reinterpret_cast<const float4*>(smem_tc + WS_SMEM_A(s1) + rec_off1);
            }
            {
                uint64_t acc01 = 0ull, acc23 = 0ull;
                tc_consume_lines<HALF>(rec0, t0, acc01, acc23);
                if (more) tc_issue_lines<HALF>(grid, R, __float_as_uint(nrec0.x), c, t0);
                float4 acc;
                f2_unpack(acc01, acc.x, acc.y); f2_unpack(acc23, acc.z, acc.w);
#ifndef NGLOD_WS_FASTSPLIT
#define NGLOD_WS_FASTSPLIT 1
#endif
#if NGLOD_WS_FASTSPLIT
#define WS_STORE_SPLIT tc_store_split4_finite
#else
#define WS_STORE_SPLIT tc_store_split4
#endif
                WS_STORE_SPLIT(a_hi, a_lo, st_off0, acc);
            }
            {
                uint64_t acc01 = 0ull, acc23 = 0ull;
                tc_consume_lines<HALF>(rec1, t1, acc01, acc23);
                if (more) tc_issue_lines<HALF>(grid, R, __float_as_uint(nrec1.x), c, t1);
                float4 acc;
                f2_unpack(acc01, acc.x, acc.y); f2_unpack(acc23, acc.z, acc.w);
                WS_STORE_SPLIT(a_hi, a_lo, st_off1, acc);
            }
            rec0 = nrec0; rec1 = nrec1;
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();                                           // release: this warp's A rows
                const unsigned old = atomicAdd(arrive_cnt(s), 1u);               // monotonic: 16 arrivals per use of the stage
                if ((old & (WS_PRODUCERS - 1)) == WS_PRODUCERS - 1) {
                    __threadfence_block();                                       // acquire: everybody's A rows
                    tc_fence_after_sync();
                    const uint32_t a_hi_s = smem_u32(a_hi);
                    tc_issue_tile(tmem_base + (uint32_t)(s * TC_N), a_hi_s, a_hi_s + TC_OPERAND_BYTES, b_hi, b_lo);
                    tc_commit(done_bar(s));
                }
            }
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ service: epilogue of tile T, set-up of tile T+4
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(WS_REGS_SERVICE));
        const int wg = (warp - WS_PRODUCERS) >> 2;        // serves tiles T = wg (mod 2)
        const int ew = warp & 3;                          // TMEM lane quarter = rows 32 ew .. 32 ew + 31
        const int row = ew * 32 + lane;
        const int R = net.res[0];
        const float* w1 = reinterpret_cast<const float*>(smem_tc + WS_SMEM_W1);
        const uint32_t rec_off = tc_elem_offset(row, WS_REC_COL), xyz_off = tc_elem_offset(row, NGLOD_F);
        const long long i_first = (long long)blockIdx.x * TC_TILE_ROWS + row;
        auto load_xyz = [&](int T, float& px, float& py, float& pz) {
            const long long i = i_first + (long long)T * tile_stride;
            px = py = pz = 0.f;
            if (T < ntiles && i < n) { px = __ldg(x + 3 * i); py = __ldg(x + 3 * i + 1); pz = __ldg(x + 3 * i + 2); }
        };
        auto setup = [&](int T, float px, float py, float pz) {
            const int s = T & (WS_STAGES - 1);
            char* a_hi = smem_tc + WS_SMEM_A(s);
            char* a_lo = a_hi + TC_OPERAND_BYTES;
            const long long i = i_first + (long long)T * tile_stride;
            // rows past n carry record 0 (corner 0, weights 0): harmless loads, their accumulator rows are never read
            const float4 rec = i < n ? tc_setup_record<HALF>(px, py, pz, R) : make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(a_hi + rec_off) = rec;
            tc_store_split4(a_hi, a_lo, xyz_off, make_float4(px, py, pz, 1.f));
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(rec_bar(s));
        };
        float ax, ay, az, bx, by, bz;                     // coordinates of this warpgroup's next two tiles (DRAM latency)
        load_xyz(wg, ax, ay, az);
        load_xyz(wg + 2, bx, by, bz);
        if (wg < ntiles) setup(wg, ax, ay, az);
        if (wg + 2 < ntiles) setup(wg + 2, bx, by, bz);
        load_xyz(wg + 4, ax, ay, az);
        for (int T = wg; T < ntiles; T += 2) {
            const int s = T & (WS_STAGES - 1), k = T / WS_STAGES;
            load_xyz(T + 6, bx, by, bz);
            mbar_wait(done_bar(s), (uint32_t)(k & 1));
            tc_fence_after_sync();
#ifndef NGLOD_WS_EPI
#define NGLOD_WS_EPI 16
#endif
#if NGLOD_WS_EPI == 8
            const float d = tc_epilogue_pipelined8(tmem_base + (uint32_t)(s * TC_N) + ((uint32_t)(ew * 32) << 16), w1);
#elif NGLOD_WS_EPI == 16
            const float d = tc_epilogue_pipelined16(tmem_base + (uint32_t)(s * TC_N) + ((uint32_t)(ew * 32) << 16), w1);
#else
            const float d = tc_epilogue(tmem_base + (uint32_t)(s * TC_N) + ((uint32_t)(ew * 32) << 16), w1);
#endif
            tc_fence_before_sync();
            const long long i = i_first + (long long)T * tile_stride;
            if (i < n) out[i] = d;
            // the stage and the accumulator of tile T are free now: set up tile T+4 in them
            if (T + 4 < ntiles) setup(T + 4, ax, ay, az);
            ax = bx; ay = by; az = bz;
        }
    }
#ifdef NGLOD_WS_TIMING      // experiment: when does every CTA finish?  (profiles/exp_ws_spread.py)
    if (threadIdx.x == 0) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
        unsigned smid;
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        printf("WSCTA %d sm %u tiles %d end_ns %llu\n", (int)blockIdx.x, smid, ntiles, t1);
    }
#endif
    tc_epilogue_free(tmem_base);
}

template <int MODE>
int launch_fwd_ws(const NetDev& nd, const float* x, long long n, float* out, cudaStream_t st) {
    auto kern = sdf_forward_ws_kernel<MODE>;
    NGLOD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM_BYTES));
    long long grid = nglod_sm_count();
    const long long want = (n + TC_TILE_ROWS - 1) / TC_TILE_ROWS;
    if (want < grid) grid = want;
    kern<<<(int)grid, WS_THREADS, WS_SMEM_BYTES, st>>>(nd, x, n, out);
    return (int)cudaGetLastError();
}

// Debug / self-test: D[128,128] = A[128,40] * B[128,40]^T through the exact operand layout, descriptors,
// 3xTF32 issue sequence and TMEM read-back the SDF kernels use.
__global__ void __launch_bounds__(128, 1)
tc_gemm_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
    extern __shared__ __align__(128) char smem_tc[];
    char* b_hi = smem_tc; char* b_lo = smem_tc + TC_OPERAND_BYTES;
    char* a_hi = smem_tc + 2 * TC_OPERAND_BYTES; char* a_lo = a_hi + TC_OPERAND_BYTES;
    char* misc = smem_tc + 4 * TC_OPERAND_BYTES;
    for (int e = threadIdx.x; e < 128 * TC_K; e += 128) {
        const int r = e / TC_K, k = e - r * TC_K;
        const uint32_t off = tc_elem_offset(r, k);
        float v = A[e], h = tf32_hi(v);
        *reinterpret_cast<float*>(a_hi + off) = h; *reinterpret_cast<float*>(a_lo + off) = v - h;
        v = B[e]; h = tf32_hi(v);
        *reinterpret_cast<float*>(b_hi + off) = h; *reinterpret_cast<float*>(b_lo + off) = v - h;
    }
    const uint32_t mbar = smem_u32(misc);
    if (threadIdx.x == 0) { mbar_init(mbar, 1); mbar_fence_init(); }
    if (threadIdx.x < 32) tmem_alloc(smem_u32(misc + 8), 128);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(misc + 8);
    if (threadIdx.x == 0) {
        tc_issue_tile(tmem_base, smem_u32(a_hi), smem_u32(a_lo), smem_u32(b_hi), smem_u32(b_lo));
        tc_commit(mbar);
    }
    mbar_wait(mbar, 0);
    tc_fence_after_sync();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int cb = 0; cb < 4; ++cb) {
        float v[32];
        tmem_ld32(trow + cb * 32, v);
        for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * 128 + cb * 32 + j] = v[j];
    }
    tc_fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem_base, 128);
}

}  // namespace

int nglod_launch_sdf_forward_tc(const NetDev& nd, const float* x, long long n, float* out, cudaStream_t st) {
#ifndef NGLOD_FWD_WS
#define NGLOD_FWD_WS 1          // 0: the serial-group kernel for the single-grid gathers too (A/B experiments)
#endif
#if NGLOD_FWD_WS
    if (nd.half_pairs) return launch_fwd_ws<TC_SINGLE_HALF>(nd, x, n, out, st);
    if (nd.num_lods == 1) return launch_fwd_ws<TC_SINGLE_F32>(nd, x, n, out, st);
#endif
    if (nd.half_pairs) return launch_fwd<TC_SINGLE_HALF>(nd, x, n, out, st);
    if (nd.num_lods == 1) return launch_fwd<TC_SINGLE_F32>(nd, x, n, out, st);
    return launch_fwd<TC_MULTI>(nd, x, n, out, st);
}

extern "C" int nglod_debug_tc_gemm(const float* A, const float* B, float* D, void* stream) {
    if (!A || !B || !D) return NGLOD_EINVAL;
    const int smem = 4 * TC_OPERAND_BYTES + 64;
    NGLOD_CUDA_TRY(cudaFuncSetAttribute(tc_gemm_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    tc_gemm_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D);
    return (int)cudaGetLastError();
}
