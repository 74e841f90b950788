// nglod_b200 -- ray vs centred unit cube (slab test of Majercik et al. 2018).
// Behavioural spec: sdf-net/lib/extensions/sol_nglod/sol_nglod_kernel.cu:93-149.
// The float operation order (double-precision reciprocal rounded to float,
// explicit fmaf for the cross terms and the entry point, first-true-of-x,y,z
// selection, the `d < 500` cut and "origin inside -> untouched defaults") is
// part of the contract: results are bit-identical to the reference kernel.
#pragma once
#include <cuda_runtime.h>

struct AabbResult {
    float x, y, z;   // entry point (ray origin when !hit)
    float t;         // entry distance (0 when !hit)
    bool hit;
};

__device__ __forceinline__ AabbResult ray_unit_cube(float ox, float oy, float oz, float dx, float dy, float dz) {
    AabbResult r;
    r.x = ox; r.y = oy; r.z = oz; r.t = 0.f; r.hit = false;
    const float cmax = fmaxf(fmaxf(fabsf(ox), fabsf(oy)), fabsf(oz));
    if (cmax < 1.0f) return r;                       // origin strictly inside: reference leaves defaults
    const float ix = (float)(1.0 / (double)dx);
    const float iy = (float)(1.0 / (double)dy);
    const float iz = (float)(1.0 / (double)dz);
    const float sx = signbit(dx) ? 1.0f : -1.0f;
    const float sy = signbit(dy) ? 1.0f : -1.0f;
    const float sz = signbit(dz) ? 1.0f : -1.0f;
    const float d0 = __fmul_rn(__fsub_rn(sx, ox), ix);
    const float d1 = __fmul_rn(__fsub_rn(sy, oy), iy);
    const float d2 = __fmul_rn(__fsub_rn(sz, oz), iz);
    const float ltxy = fmaf(dy, d0, oy), ltxz = fmaf(dz, d0, oz);
    const float ltyx = fmaf(dx, d1, ox), ltyz = fmaf(dz, d1, oz);
    const float ltzx = fmaf(dx, d2, ox), ltzy = fmaf(dy, d2, oy);
    const bool t0 = (d0 >= 0.0f) && (fabsf(ltxy) < 1.0f) && (fabsf(ltxz) < 1.0f);
    const bool t1 = (d1 >= 0.0f) && (fabsf(ltyx) < 1.0f) && (fabsf(ltyz) < 1.0f);
    const bool t2 = (d2 >= 0.0f) && (fabsf(ltzx) < 1.0f) && (fabsf(ltzy) < 1.0f);
    float d = 0.0f;
    bool any = true;
    if (t0) d = d0; else if (t1) d = d1; else if (t2) d = d2; else any = false;
    if (any && d < 500.0f) {
        r.t = d; r.hit = true;
        r.x = fmaf(dx, d, ox); r.y = fmaf(dy, d, oy); r.z = fmaf(dz, d, oz);
    }
    return r;
}
