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

/* ------------------------------------------------------------------------------------------------
 * SPC (sparse octree) ray traversal and first-voxel search.
 *
 *   oracle_spc_raytrace  <- sol-renderer/include/spc/spc/spc_raytrace_cuda_kernel.cu:85-265
 *                           (d_Decide :108-138, d_FaceEval :85-105, d_Subdivide :141-180, d_Compactify :183-190)
 *   oracle_spc_ray_aabb  <- sol-renderer/include/solr/solr/gfx/ray_aabb.cuh:42-192
 *
 * The reference runs level-synchronously over all (ray, voxel) "nuggets": Decide -> inclusive scan -> Subdivide
 * writes each surviving nugget's occupied children, in front-to-back order, at its scanned offset.  Expanding
 * every nugget IN PLACE level after level yields, per ray, exactly the pre-order leaf sequence of a depth-first
 * walk with the same child order, and rays stay in index order -- which is what is computed here, ray by ray.
 *
 * Parity status: UNPINNED.  These kernels exist only as CUDA, and they cannot be compiled here (they need
 * CUDA-samples' helper_math.h and a pre-2.0 CUB, see oracle/build_ref.py).  Float expression shapes follow what
 * nvcc 12.9 emits for the reference's source (checked in SASS): o = fma(0.5,org,0.5); cross terms
 * fma(a.y,b.z,-(b.y*a.z)); FaceEval r0 = fma(b,j,a*i) + c.
 * ------------------------------------------------------------------------------------------------ */
static const unsigned char SPC_ORDER[8][8] = {          /* spc_raytrace_cuda_kernel.cu:39-47 */
    {0, 1, 2, 4, 3, 5, 6, 7}, {1, 0, 3, 5, 2, 4, 7, 6}, {2, 0, 3, 6, 1, 4, 7, 5}, {3, 1, 2, 7, 0, 5, 6, 4},
    {4, 0, 5, 6, 1, 2, 7, 3}, {5, 1, 4, 7, 0, 3, 6, 2}, {6, 2, 4, 7, 0, 3, 5, 1}, {7, 3, 5, 6, 1, 2, 4, 0}};

static int spc_face_eval(unsigned short i, unsigned short j, float a, float b, float c) {
    float r[4];
    r[0] = fmaf(b, (float)j, a * (float)i) + c;
    r[1] = r[0] + a;
    r[2] = r[0] + b;
    r[3] = r[1] + b;
    float mn = 1.0f, mx = -1.0f;
    for (int k = 0; k < 4; ++k) { if (r[k] < mn) mn = r[k]; if (r[k] > mx) mx = r[k]; }
    return mn <= 0.0f && mx >= 0.0f;
}

static int spc_decide(const short* p, const float* org, const float* dir, int level) {
    const float ox = fmaf(0.5f, org[0], 0.5f), oy = fmaf(0.5f, org[1], 0.5f), oz = fmaf(0.5f, org[2], 0.5f);
    const float dx = 0.5f * dir[0], dy = 0.5f * dir[1], dz = 0.5f * dir[2];
    const float cx = fmaf(oy, dz, -(dy * oz)), cy = fmaf(oz, dx, -(dz * ox)), cz = fmaf(ox, dy, -(dx * oy));
    const float s1 = 1.0f / (float)(1 << level), s2 = s1 * s1;
    return spc_face_eval((unsigned short)p[1], (unsigned short)p[2], -s2 * dz, s2 * dy, s1 * cx) &&
           spc_face_eval((unsigned short)p[0], (unsigned short)p[2], s2 * dz, -s2 * dx, s1 * cy) &&
           spc_face_eval((unsigned short)p[0], (unsigned short)p[1], -s2 * dy, s2 * dx, s1 * cz);
}

typedef struct {
    const uint8_t* octree; const int32_t* prefix; const short* points; const int32_t* pyrsum;
    int target; const float* org; const float* dir; int32_t ray; int32_t* out; int64_t cap; int64_t count;
} spc_ctx;

static void spc_visit(spc_ctx* c, int level, int32_t pidx) {
    const int32_t g = c->pyrsum[level] + pidx;
    const short* p = c->points + 4 * (int64_t)g;
    if (!spc_decide(p, c->org, c->dir, level)) return;
    if (level == c->target) {
        if (c->out && c->count < c->cap) { c->out[2 * c->count] = c->ray; c->out[2 * c->count + 1] = pidx; }
        c->count++;
        return;
    }
    const unsigned o = c->octree[g];
    const int32_t s = c->prefix[g];
    const float scale = 1.0f / (float)(1 << level);
    /* octant of the ray origin relative to the voxel centre (exact in double, :160-167) */
    const double x = (double)fmaf(0.5f, c->org[0], 0.5f) - (double)scale * ((double)p[0] + 0.5);
    const double y = (double)fmaf(0.5f, c->org[1], 0.5f) - (double)scale * ((double)p[1] + 0.5);
    const double z = (double)fmaf(0.5f, c->org[2], 0.5f) - (double)scale * ((double)p[2] + 0.5);
    int code = 0;
    if ((float)x > 0) code = 4;
    if ((float)y > 0) code += 2;
    if ((float)z > 0) code += 1;
    for (int i = 0; i < 8; ++i) {
        const int j = SPC_ORDER[code][i];
        if (o & (1u << j)) {
            const int cnt = __builtin_popcount(o & ((2u << j) - 1u));
            spc_visit(c, level + 1, s + cnt - c->pyrsum[level + 1]);
        }
    }
}

/* octree: child masks (breadth first); prefix: exclusive sum of popcounts; points: [psize,4] int16 per level in
 * Morton order; pyrsum[l]: first point of level l.  Writes up to `cap` nuggets (ray, point index within the target
 * level); ray_count[r] (optional) = nuggets of ray r.  Returns the total. */
int64_t oracle_spc_raytrace(const uint8_t* octree, const int32_t* prefix, const short* points, const int32_t* pyrsum,
                            int target, const float* ray_o, const float* ray_d, int64_t n, int32_t* nuggets,
                            int64_t cap, int32_t* ray_count) {
    spc_ctx c = {octree, prefix, points, pyrsum, target, 0, 0, 0, nuggets, cap, 0};
    for (int64_t r = 0; r < n; ++r) {
        const int64_t before = c.count;
        c.org = ray_o + 3 * r; c.dir = ray_d + 3 * r; c.ray = (int32_t)r;
        spc_visit(&c, 0, 0);
        if (ray_count) ray_count[r] = (int32_t)(c.count - before);
    }
    return c.count;
}

static float spc_ray_aabb1(const float* q, const float* dir, const float* inv, const float* sgn, const float* vc, float r) {
    const float ox = q[0] - vc[0], oy = q[1] - vc[1], oz = q[2] - vc[2];          /* ray_aabb.cuh:52-55 */
    const float cmax = fmaxf(fmaxf(fabsf(ox), fabsf(oy)), fabsf(oz));
    float winding = cmax < r ? -1.0f : 1.0f;
    winding *= r;
    if (winding < 0) return winding;
    const float d0 = fmaf(winding, sgn[0], -ox) * inv[0], d1 = fmaf(winding, sgn[1], -oy) * inv[1],
                d2 = fmaf(winding, sgn[2], -oz) * inv[2];
    const float ltxy = fmaf(dir[1], d0, oy), ltxz = fmaf(dir[2], d0, oz);
    const float ltyx = fmaf(dir[0], d1, ox), ltyz = fmaf(dir[2], d1, oz);
    const float ltzx = fmaf(dir[0], d2, ox), ltzy = fmaf(dir[1], d2, oy);
    const int t0 = (d0 >= 0.0f) && (fabsf(ltxy) < r) && (fabsf(ltxz) < r);
    const int t1 = (d1 >= 0.0f) && (fabsf(ltyx) < r) && (fabsf(ltyz) < r);
    const int t2 = (d2 >= 0.0f) && (fabsf(ltzx) < r) && (fabsf(ltzy) < r);
    if (t0) return d0;
    if (t1) return d1;
    if (t2) return d2;
    return 0.0f;
}

/* ray_aabb_kernel (nugget version, ray_aabb.cuh:104-192) with init = true and query = ray origin + t*dir given
 * by the caller in `query`: for every ray that owns a nugget run, the first voxel of the run that contains the
 * query point (d = -r, no advance) or that the ray enters (d > 0, t += d).  Rays without a run are left untouched.
 * level_points: points of the nugget level [*,4] int16. */
void oracle_spc_ray_aabb(const int32_t* nuggets, int64_t num_nuggets, const short* level_points, int level,
                         const float* ray_o, const float* ray_d, const float* query, const uint8_t* active,
                         float* x, float* t, uint8_t* cond, int32_t* pidx) {
    const float r = 1.0f / (float)(1 << level);
    int64_t i = 0;
    while (i < num_nuggets) {
        const int32_t ray = nuggets[2 * i];
        if (active && !active[ray]) {                       /* `if (!cond[ridx] && !init) continue;` (:130) */
            while (i < num_nuggets && nuggets[2 * i] == ray) ++i;
            continue;
        }
        const float* dir = ray_d + 3 * (int64_t)ray;
        const float inv[3] = {1.0f / dir[0], 1.0f / dir[1], 1.0f / dir[2]};
        const float sgn[3] = {signbit(dir[0]) ? 1.0f : -1.0f, signbit(dir[1]) ? 1.0f : -1.0f, signbit(dir[2]) ? 1.0f : -1.0f};
        int hit = 0;
        int64_t j = i;
        for (; j < num_nuggets && nuggets[2 * j] == ray; ++j) {
            if (hit) continue;
            const short* p = level_points + 4 * (int64_t)nuggets[2 * j + 1];
            const float vc[3] = {fmaf(r, fmaf(2.0f, (float)p[0], 1.0f), -1.0f), fmaf(r, fmaf(2.0f, (float)p[1], 1.0f), -1.0f),
                                 fmaf(r, fmaf(2.0f, (float)p[2], 1.0f), -1.0f)};
            const float d = spc_ray_aabb1(query + 3 * (int64_t)ray, dir, inv, sgn, vc, r);
            if (d != 0.0f) {
                hit = 1;
                pidx[ray] = nuggets[2 * j + 1];
                cond[ray] = 1;
                if (d > 0.0f) {
                    t[ray] += d;
                    for (int k = 0; k < 3; ++k) x[3 * (int64_t)ray + k] = fmaf(dir[k], t[ray], ray_o[3 * (int64_t)ray + k]);
                }
            }
        }
        if (!hit) {
            cond[ray] = 0;
            t[ray] = 100.0f;
            for (int k = 0; k < 3; ++k) x[3 * (int64_t)ray + k] = fmaf(dir[k], 100.0f, ray_o[3 * (int64_t)ray + k]);
        }
        i = j;
    }
}
