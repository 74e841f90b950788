// nglod_b200 -- hand-written tcgen05 / TMEM / mbarrier plumbing for sm_100a (inline PTX, no CUTLASS).
//
// The decoder's 35->128 contraction runs on the 5th-gen tensor cores as
//     D[128 points x 128 hidden] (fp32, TMEM)  =  A[128 x 40] * B[128 x 40]^T       (kind::tf32)
// with A = {32 interpolated features, x, y, z, 1, 0,0,0,0} produced by threads into shared memory and
// B = {W0 feature cols, W0 xyz cols, b0, 0,0,0,0} staged once per CTA.  To keep FP32-level accuracy
// (needed by the tracer's finite-difference normals) every operand is split
// x = hi + lo with hi = rna-rounded TF32, and D accumulates  A_lo*B_hi + A_hi*B_lo + A_hi*B_hi  (3xTF32).
//
// Shared-memory operand layout (K-major, SWIZZLE_NONE "interleaved" canonical layout, in bytes):
//     elem(row, k) at  (row/8)*SBO + (k/4)*LBO + (row%8)*16 + (k%4)*4
// i.e. a core matrix = 8 rows x 16 bytes, contiguous (128 B).  LBO is padded to 144 B (not 128) so that the
// 8 lanes of a query, which each store one 16-byte K-chunk of the SAME row, land in 8 different bank groups.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define TC_K 40                         // padded contraction length (5 MMA k-steps of 8)
#define TC_KCHUNKS (TC_K / 4)           // 16-byte chunks per row
#define TC_LBO 144u                     // bytes between consecutive K chunks (128 + 16 pad)
#define TC_SBO (TC_KCHUNKS * TC_LBO)    // bytes between 8-row groups = 1440
#define TC_TILE_ROWS 128
#define TC_OPERAND_BYTES ((TC_TILE_ROWS / 8) * TC_SBO)   // 23040 B per 128-row operand
#define TC_N 128                        // hidden units = MMA N
// instruction descriptor, kind::tf32: D=f32 (1<<4), A=B=tf32 (2<<7, 2<<10), K-major both, N>>3 at bit 17, M>>4 at bit 24
#define TC_IDESC ((1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24))

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t tc_elem_offset(int row, int k) {
    return (uint32_t)(row >> 3) * TC_SBO + (uint32_t)(k >> 2) * TC_LBO + (uint32_t)(row & 7) * 16u + (uint32_t)(k & 3) * 4u;
}

// 64-bit shared-memory matrix descriptor (SWIZZLE_NONE, version 1)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr) {
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((TC_LBO >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((TC_SBO >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;
    return d;
}

__device__ __forceinline__ float tf32_hi(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// cvt.rna.tf32.f32 for FINITE inputs as two integer instructions (round the magnitude to nearest, ties away, at bit 13):
// ptxas expands the PTX conversion into five (it guards Inf / NaN, which interpolated features never are).
__device__ __forceinline__ float tf32_hi_finite(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(TC_IDESC), "r"(accumulate) : "memory");
}

// D = A_lo*B_hi + A_hi*B_lo + A_hi*B_hi over 5 k-steps each (15 MMAs), issued by ONE thread.
__device__ __forceinline__ void tc_issue_tile(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo) {
    uint32_t acc = 0;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a = pass == 0 ? a_lo : a_hi;
        const uint32_t b = pass == 1 ? b_lo : b_hi;
#pragma unroll
        for (int ks = 0; ks < TC_K / 8; ++ks) {
            tc_mma_tf32(d_tmem, tc_smem_desc(a + ks * 2 * TC_LBO), tc_smem_desc(b + ks * 2 * TC_LBO), acc);
            acc = 1;
        }
    }
}

// One lane of a CONVERGED warp (the same one every time).  tcgen05.mma takes its descriptors from uniform registers: issued
// from a lane-divergent branch (`if (lane == 0)`) every descriptor is computed in vector registers and moved over with
// R2UR; with the whole warp in the branch and only the instruction itself under the elected predicate the descriptor
// arithmetic stays in the uniform datapath (SASS: UMOV / ULOP3 / UIADD3 between back-to-back UTCHMMA).  Measured: it buys
// nothing -- the issuing thread is paced by the tensor pipe accepting the instruction (~80-90 cycles per MMA of a chain
// into one accumulator), not by its operand arithmetic: neutral in the backward kernel (which uses it), a LOSS in the
// forward (124 -> 133 us) and the tracer (0.90 -> 0.96 ms per frame), where the 31 idle lanes of the issuing warp delay
// that warp's own gather / epilogue work.
__device__ __forceinline__ bool tc_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// v is the same in every lane: tell the compiler (values derived from it can live in uniform registers)
__device__ __forceinline__ int tc_warp_uniform(int v) { return __shfl_sync(0xffffffffu, v, 0); }

// tc_issue_tile for a converged warp: every lane calls, one elected lane issues.
__device__ __forceinline__ void tc_issue_tile_warp(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo) {
    uint32_t acc = 0;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a = pass == 0 ? a_lo : a_hi;
        const uint32_t b = pass == 1 ? b_lo : b_hi;
#pragma unroll
        for (int ks = 0; ks < TC_K / 8; ++ks) {
            const uint64_t ad = tc_smem_desc(a + ks * 2 * TC_LBO), bd = tc_smem_desc(b + ks * 2 * TC_LBO);
            if (tc_elect_one()) tc_mma_tf32(d_tmem, ad, bd, acc);
            acc = 1;
        }
    }
}

__device__ __forceinline__ void tc_commit(uint32_t mbar_saddr) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mbar_saddr) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint32_t saddr, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(saddr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t saddr, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "MBAR_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra MBAR_DONE;\n\t"
        "bra MBAR_WAIT;\n\t"
        "MBAR_DONE:\n\t}"
        :: "r"(saddr), "r"(parity) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t saddr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(saddr) : "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
}

// TMEM allocation: one full warp, power-of-two columns >= 32; base address lands in shared memory.
__device__ __forceinline__ void tmem_alloc(uint32_t dst_saddr, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(dst_saddr), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}

// 16 consecutive fp32 columns of this thread's TMEM lane, WITHOUT waiting: pair with tmem_ld_wait() before the first use.
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld8_async(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// This thread's TMEM lane (= accumulator row), 32 consecutive fp32 columns starting at taddr's column.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
