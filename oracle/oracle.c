/*
 * oracle.c -- CPU restatement of the reference's NATIVE kernels on the hot path.
 *
 * TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg as the CHECKER; never by the product path (nglod_b200/).
 *
 * Parity status: the reference ships no tests or golden vectors for these kernels
 *  and both exist only as CUDA, so in the authoring container this
 * restatement is "parity unpinned"; it is pinned on the GPU box, where the `-m gpu` tests
 * compare BOTH this file and the sm_100a kernels against the reference's own kernels compiled
 * unmodified into oracle/_ref/ (oracle/build_ref.py).
 *
 *   oracle_aabb      <- sdf-net/lib/extensions/sol_nglod/sol_nglod_kernel.cu:93-149
 *   oracle_mesh2sdf  <- sdf-net/lib/extensions/mesh2sdf_cuda/mesh2sdf_kernel.cu:307-616
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC oracle.c -o liboracle.so -lm
 * (contraction off: every fused multiply-add below is an explicit fmaf, as in the reference).
 */
#include <math.h>
#include <stdint.h>

/* sol_nglod_kernel.cu:93-149, one ray at a time. */
void oracle_aabb(const float* ray_o, const float* ray_d, int64_t n, float* x, float* t, uint8_t* hit) {
    for (int64_t i = 0; i < n; ++i) {
        const float ox = ray_o[3 * i], oy = ray_o[3 * i + 1], oz = ray_o[3 * i + 2];
        const float dx = ray_d[3 * i], dy = ray_d[3 * i + 1], dz = ray_d[3 * i + 2];
        /* defaults set by the wrapper (:167-173): x = clone(ray_o), t = 0, hit = false */
        x[3 * i] = ox; x[3 * i + 1] = oy; x[3 * i + 2] = oz; t[i] = 0.0f; hit[i] = 0;
        /* :94  1.0/ray_d is evaluated in double and rounded into a float3 */
        const float ix = (float)(1.0 / (double)dx), iy = (float)(1.0 / (double)dy), iz = (float)(1.0 / (double)dz);
        const float s0 = signbit(dx) ? 1.0f : -1.0f, s1 = signbit(dy) ? 1.0f : -1.0f, s2 = signbit(dz) ? 1.0f : -1.0f;
        const float cmax = fmaxf(fmaxf(fabsf(ox), fabsf(oy)), fabsf(oz));
        if (cmax < 1.0f) continue;                                   /* :108-112 origin inside */
        const float d0 = (s0 - ox) * ix, d1 = (s1 - oy) * iy, d2 = (s2 - oz) * iz;   /* :114-116 */
        const float ltxy = fmaf(dy, d0, oy), ltxz = fmaf(dz, d0, oz);                /* :118-125 */
        const float ltyx = fmaf(dx, d1, ox), ltyz = fmaf(dz, d1, oz);
        const float ltzx = fmaf(dx, d2, ox), ltzy = fmaf(dy, d2, oy);
        const int t0 = (d0 >= 0.0f) && (fabsf(ltxy) < 1.0f) && (fabsf(ltxz) < 1.0f); /* :127-129 */
        const int t1 = (d1 >= 0.0f) && (fabsf(ltyx) < 1.0f) && (fabsf(ltyz) < 1.0f);
        const int t2 = (d2 >= 0.0f) && (fabsf(ltzx) < 1.0f) && (fabsf(ltzy) < 1.0f);
        float d; int any = 1;
        if (t0) d = d0; else if (t1) d = d1; else if (t2) d = d2; else { d = 0.0f; any = 0; }   /* :131-140 */
        if (any && d < 500.0f) {                                     /* :142-149 */
            t[i] = d; hit[i] = 1;
            x[3 * i] = fmaf(dx, d, ox); x[3 * i + 1] = fmaf(dy, d, oy); x[3 * i + 2] = fmaf(dz, d, oz);
        }
    }
}

static float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void cross3(const float* a, const float* b, float* r) {
    r[0] = a[1] * b[2] - a[2] * b[1]; r[1] = a[2] * b[0] - a[0] * b[2]; r[2] = a[0] * b[1] - a[1] * b[0];
}
static float idot2(const float* a) { return 1.0f / (a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); } /* __frcp_rn */
static float clampf(float f, float a, float b) { return fmaxf(a, fminf(f, b)); }
static float d2axmb(const float* a, float x, const float* b) {
    const float t0 = a[0] * x - b[0], t1 = a[1] * x - b[1], t2 = a[2] * x - b[2];
    return t0 * t0 + t1 * t1 + t2 * t2;
}

static const float STAB[13][3] = {                                   /* mesh2sdf_kernel.cu:341-353 */
    {1.0f, 0.0f, 0.0f}, {0.0f, 1.0f, 0.0f}, {0.0f, 0.0f, 1.0f},
    {0.0f, 0.707106781f, 0.707106781f}, {0.707106781f, 0.0f, 0.707106781f}, {0.707106781f, 0.707106781f, 0.0f},
    {0.0f, 0.707106781f, -0.707106781f}, {0.707106781f, 0.0f, -0.707106781f}, {0.707106781f, -0.707106781f, 0.0f},
    {0.577350269f, 0.577350269f, 0.577350269f}, {-0.577350269f, 0.577350269f, 0.577350269f},
    {0.577350269f, -0.577350269f, 0.577350269f}, {0.577350269f, 0.577350269f, -0.577350269f}};

/* kernel_mesh2sdf_quad (:307-556) + kernel_quad_aggr (:558-616); the 64-way triangle split of the
 * reference only changes the order of min / OR reductions, which are order-independent. */
void oracle_mesh2sdf(const float* points, int64_t n, const float* mesh, int64_t nt, float* out) {
    for (int64_t p = 0; p < n; ++p) {
        const float* P = points + 3 * p;
        float mind2 = INFINITY;
        int pos[13] = {0}, neg[13] = {0};
        for (int64_t i = 0; i < nt; ++i) {
            const float* T = mesh + 9 * i;
            float v10[3], v21[3], v02[3], nor[3], c10[3], c21[3], c02[3], p0[3], p1[3], p2[3];
            for (int k = 0; k < 3; ++k) {
                v10[k] = T[3 + k] - T[k]; v21[k] = T[6 + k] - T[3 + k]; v02[k] = T[k] - T[6 + k];
                p0[k] = P[k] - T[k]; p1[k] = P[k] - T[3 + k]; p2[k] = P[k] - T[6 + k];
            }
            cross3(v10, v02, nor);
            cross3(v10, nor, c10); cross3(v21, nor, c21); cross3(v02, nor, c02);
            if (nor[0] != 0.0f || nor[1] != 0.0f || nor[2] != 0.0f) {            /* :422 */
                const float s1 = copysignf(1.0f, dot3(c10, p0)), s2 = copysignf(1.0f, dot3(c21, p1)),
                            s3 = copysignf(1.0f, dot3(c02, p2));
                float d2;
                if ((s1 + s2 + s3) < 2.0f) {                                     /* :430 edge distance */
                    const float e1 = d2axmb(v10, clampf(dot3(v10, p0) * idot2(v10), 0.0f, 1.0f), p0);
                    const float e2 = d2axmb(v21, clampf(dot3(v21, p1) * idot2(v21), 0.0f, 1.0f), p1);
                    const float e3 = d2axmb(v02, clampf(dot3(v02, p2) * idot2(v02), 0.0f, 1.0f), p2);
                    d2 = fminf(e1, fminf(e2, e3));
                } else {                                                         /* :437 face distance */
                    d2 = dot3(nor, p0) * dot3(nor, p0) * idot2(nor);
                }
                if (d2 < 0.0f) d2 = 0.0f;
                mind2 = fminf(mind2, d2);
            }
            const float edge2[3] = {-v02[0], -v02[1], -v02[2]};
            for (int k = 0; k < 13; ++k) {                                       /* :479-526 */
                float pvec[3], qvec[3];
                cross3(STAB[k], edge2, pvec);
                const float det = dot3(v10, pvec);
                if (det > -1e-8 && det < 1e-8) continue;
                const float inv_det = 1.0f / det;
                const float u = dot3(p0, pvec) * inv_det;
                if (u < 0.0f || u > 1.0f) continue;
                cross3(p0, v10, qvec);
                const float v = dot3(STAB[k], qvec) * inv_det;
                if (v < 0.0f || u + v > 1.0f) continue;
                const float tt = dot3(edge2, qvec) * inv_det;
                if (tt >= 0.0f) pos[k] = 1; else neg[k] = 1;
            }
        }
        int outside = 0;
        for (int k = 0; k < 13; ++k) if (!pos[k] || !neg[k]) { outside = 1; break; }   /* :573-589 */
        if (mind2 < 0.0f) mind2 = 0.0f;
        float d = sqrtf(mind2);
        out[p] = outside ? d : -d;
    }
}
