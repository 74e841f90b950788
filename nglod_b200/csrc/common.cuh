// nglod_b200 -- shared device/host helpers (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/nglod_b200.h"

#define NGLOD_F 32        // feature channels the fast kernels are built for
#define NGLOD_H 128       // decoder hidden width the fast kernels are built for
#define NGLOD_KPAD 36     // decoder input padded: 32 feat + xyz + 1.0 (bias column)

#define NGLOD_CUDA_TRY(expr)                        \
    do {                                            \
        cudaError_t _e = (expr);                    \
        if (_e != cudaSuccess) return (int)_e;      \
    } while (0)

// Device-side view of the model for ONE decoder head (kernel parameter, by value).
struct NetDev {
    int num_lods;       // number of grids to sum (= lod + 1)
    int pos_invariant;
    int half_pairs;     // 1: grids[0] is an fp16 x-pair-line grid (nglod_pack_grid_fp16); single-LOD inference views only
    int res[NGLOD_MAX_LODS];
    const float* grids[NGLOD_MAX_LODS];
    const float* w0;    // [H, in_dim]
    const float* b0;    // [H]
    const float* w1;    // [H]
    const float* b1;    // [1]
};

struct GradDev {
    float* grids[NGLOD_MAX_LODS];
    float* w0;
    float* b0;
    float* w1;
    float* b1;
    // tcgen05 backward only: CTA b scatters into priv + (b % priv_copies) * priv_stride instead of grids[0] when priv is set
    float* priv = nullptr;
    int priv_copies = 0;
    int priv_stride = 0;        // floats
};

static inline int nglod_check_net(const nglod_net_t* net, int lod) {
    if (!net) return NGLOD_EINVAL;
    if (net->num_lods < 1 || net->num_lods > NGLOD_MAX_LODS) return NGLOD_EINVAL;
    if (lod < 0 || lod >= net->num_lods) return NGLOD_EINVAL;
    if (net->feature_dim != NGLOD_F || net->hidden_dim != NGLOD_H) return NGLOD_EUNSUPPORTED;
    if (net->math_mode != NGLOD_MATH_TC3XTF32 && net->math_mode != NGLOD_MATH_FP32) return NGLOD_EINVAL;
    for (int i = 0; i <= lod; ++i) {
        if (!net->grids[i] || net->grid_res[i] < 1 || net->grid_res[i] > 256) return NGLOD_EINVAL;  // 32-bit element offsets
        if ((reinterpret_cast<uintptr_t>(net->grids[i]) & 15u) != 0) return NGLOD_EINVAL;
    }
    if (!net->w0[lod] || !net->b0[lod] || !net->w1[lod] || !net->b1[lod]) return NGLOD_EINVAL;
    if ((reinterpret_cast<uintptr_t>(net->summed[lod]) & 15u) || (reinterpret_cast<uintptr_t>(net->summed_fp16[lod]) & 127u))
        return NGLOD_EINVAL;
    return 0;
}

static inline NetDev nglod_make_netdev(const nglod_net_t* net, int lod) {
    NetDev d;
    d.num_lods = lod + 1;
    d.pos_invariant = net->pos_invariant;
    d.half_pairs = 0;
    for (int i = 0; i < NGLOD_MAX_LODS; ++i) {
        d.res[i] = i <= lod ? net->grid_res[i] : 1;
        d.grids[i] = i <= lod ? net->grids[i] : nullptr;
    }
    d.w0 = net->w0[lod];
    d.b0 = net->b0[lod];
    d.w1 = net->w1[lod];
    d.b1 = net->b1[lod];
    return d;
}

// View for the inference kernels: when the caller supplied the prefix-summed grid of this LOD (nglod_net_t.summed),
// gather that ONE grid instead of lod+1 of them; tensor-core mode prefers its fp16 x-pair copy when present.
static inline NetDev nglod_make_netdev_infer(const nglod_net_t* net, int lod, bool allow_half = true) {
    NetDev d = nglod_make_netdev(net, lod);
    if (!net->summed[lod]) return d;
    d.num_lods = 1;
    d.res[0] = net->grid_res[lod];
    d.grids[0] = net->summed[lod];
    for (int i = 1; i < NGLOD_MAX_LODS; ++i) { d.res[i] = 1; d.grids[i] = nullptr; }
    if (allow_half && net->math_mode == NGLOD_MATH_TC3XTF32 && net->summed_fp16[lod]) {
        d.half_pairs = 1;
        d.grids[0] = reinterpret_cast<const float*>(net->summed_fp16[lod]);
    }
    return d;
}

// Cached device properties (SM count) -- read-only after first use.
static inline int nglod_sm_count() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
        sms = v;
    }
    return sms;
}

__device__ __forceinline__ float4 ldg_f4(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}

// ---- packed fp32x2 arithmetic (sm_100a FFMA2 / FMUL2 / FADD2: two IEEE fp32 results per instruction, each lane
//      rounded exactly like the scalar op, so results are bit-identical to the scalar code they replace)
__device__ __forceinline__ uint64_t f2_pack(float a, float b) {
    uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& a, float& b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
    uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
    uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
// 128-bit vector reduction (sm_90+): one L2 RED op for 4 consecutive floats.
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
