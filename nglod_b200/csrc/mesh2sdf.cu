// nglod_b200 -- brute-force mesh -> signed distance for the training sampler.
//
// Behavioural spec: kernel_mesh2sdf_quad + kernel_quad_aggr,
// sdf-net/lib/extensions/mesh2sdf_cuda/mesh2sdf_kernel.cu:307-616 (entry :895-927):
//   unsigned distance = sqrt(min over triangles of the edge / face squared distance)
//   inside  <=>  for ALL 13 fixed stab directions the line through the point hits
//                at least one triangle at t >= 0 AND at least one at t < 0.
//
// Two paths, bit-identical to each other (tests/): the brute-force walk described next, for small batches and small
// meshes, and -- further down -- an output-sensitive pair of kernels for large batches (distance through a sphere
// hierarchy over Morton-sorted triangles, sign through 13 projected point grids).
//
// B200 design (not the reference's): one thread per point keeps its running
// min and two 13-bit stab masks in registers; triangles are streamed through
// shared memory in tiles of "records" holding everything that depends only on
// the triangle -- edge vectors, the normal, the three edge-plane normals, the
// four reciprocals and, per stab direction, pvec = dir x edge2 and 1/det --
// computed once per CTA per tile instead of once per (point, triangle).  No
// [64, N] distance / [64, N, 13, 2] flag temporaries (0.96 GB at N = 500k in the
// reference), no second aggregation kernel.  All reads of a record are warp-
// uniform shared-memory broadcasts.
//
// The per-pair arithmetic keeps the reference's expression shapes (dot products
// left to right, a*x-b residuals, IEEE 1/x) so that nvcc's FMA contraction lands
// on the same instruction sequence; the parity test compares against the
// reference's own compiled kernel on the GPU.
#include "common.cuh"
#include <mutex>

namespace {

constexpr int M2S_THREADS = 256;
constexpr int M2S_TILE = 48;          // triangles per shared-memory tile
constexpr int M2S_NDIR = 13;
constexpr int M2S_PATCH = 32;         // triangles per patch of the culling hierarchy (one lane each)
#ifndef NGLOD_M2S_UNROLL
#define NGLOD_M2S_UNROLL 1
#endif
constexpr int M2S_UNROLL = NGLOD_M2S_UNROLL;      // triangles per loop iteration of the main kernel

struct __align__(16) TriRecord {
    float a[3], b[3], c[3];           // vertices
    float v10[3], v21[3], v02[3];     // b-a, c-b, a-c
    float nor[3];                     // v10 x v02
    float c10[3], c21[3], c02[3];     // edge x nor
    float inv10, inv21, inv02, invn;  // 1/|.|^2  (round-to-nearest reciprocal)
    float pvec[M2S_NDIR][3];          // dir x edge2,  edge2 = -v02
    float inv_det[M2S_NDIR];          // 1 / (edge1 . pvec), edge1 = v10
    unsigned dir_ok;                  // bit k: |det_k| >= 1e-8 (else the reference skips the direction)
    unsigned nondegenerate;           // nor != 0
    float cen[3];                     // bounding sphere (centroid, radius with slack): distance culling only
    float rad;
    unsigned pad[4];                  // to 384 B with every word written (the tiles are staged as whole 16-byte words)
};
static_assert(sizeof(TriRecord) == 384, "record layout");

__constant__ float c_stab_dir[M2S_NDIR][3] = {
    {1.0f, 0.0f, 0.0f}, {0.0f, 1.0f, 0.0f}, {0.0f, 0.0f, 1.0f},
    {0.0f, 0.707106781f, 0.707106781f}, {0.707106781f, 0.0f, 0.707106781f}, {0.707106781f, 0.707106781f, 0.0f},
    {0.0f, 0.707106781f, -0.707106781f}, {0.707106781f, 0.0f, -0.707106781f}, {0.707106781f, -0.707106781f, 0.0f},
    {0.577350269f, 0.577350269f, 0.577350269f}, {-0.577350269f, 0.577350269f, 0.577350269f},
    {0.577350269f, -0.577350269f, 0.577350269f}, {0.577350269f, 0.577350269f, -0.577350269f}};

__device__ __forceinline__ float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ void cross3(const float* a, const float* b, float* r) {
    r[0] = a[1] * b[2] - a[2] * b[1];
    r[1] = a[2] * b[0] - a[0] * b[2];
    r[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ float rcp_dot2(const float* a) { return __frcp_rn(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
__device__ __forceinline__ float clamp01(float f) { return fmaxf(0.0f, fminf(f, 1.0f)); }
__device__ __forceinline__ float d2axmb(const float* a, float x, const float* b) {
    const float t0 = a[0] * x - b[0];
    const float t1 = a[1] * x - b[1];
    const float t2 = a[2] * x - b[2];
    return t0 * t0 + t1 * t1 + t2 * t2;
}

__device__ void build_record(const float* __restrict__ tri, TriRecord& r) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { r.a[k] = tri[k]; r.b[k] = tri[3 + k]; r.c[k] = tri[6 + k]; }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        r.v10[k] = r.b[k] - r.a[k];
        r.v21[k] = r.c[k] - r.b[k];
        r.v02[k] = r.a[k] - r.c[k];
    }
    cross3(r.v10, r.v02, r.nor);
    cross3(r.v10, r.nor, r.c10);
    cross3(r.v21, r.nor, r.c21);
    cross3(r.v02, r.nor, r.c02);
    r.inv10 = rcp_dot2(r.v10);
    r.inv21 = rcp_dot2(r.v21);
    r.inv02 = rcp_dot2(r.v02);
    r.invn = rcp_dot2(r.nor);
    r.nondegenerate = (r.nor[0] != 0.0f || r.nor[1] != 0.0f || r.nor[2] != 0.0f) ? 1u : 0u;
    float edge2[3] = {-r.v02[0], -r.v02[1], -r.v02[2]};
    unsigned ok = 0;
    for (int k = 0; k < M2S_NDIR; ++k) {
        float pv[3];
        cross3(c_stab_dir[k], edge2, pv);
        const float det = dot3(r.v10, pv);
        r.pvec[k][0] = pv[0]; r.pvec[k][1] = pv[1]; r.pvec[k][2] = pv[2];
        if (!(det > -1e-8 && det < 1e-8)) {       // note: double literals, as in the reference
            ok |= 1u << k;
            r.inv_det[k] = 1.0f / det;
        } else {
            r.inv_det[k] = 0.0f;
        }
    }
    r.dir_ok = ok;
    float r2 = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) r.cen[k] = (r.a[k] + r.b[k] + r.c[k]) * (1.0f / 3.0f);
    {
        const float da[3] = {r.a[0] - r.cen[0], r.a[1] - r.cen[1], r.a[2] - r.cen[2]};
        const float db[3] = {r.b[0] - r.cen[0], r.b[1] - r.cen[1], r.b[2] - r.cen[2]};
        const float dc[3] = {r.c[0] - r.cen[0], r.c[1] - r.cen[1], r.c[2] - r.cen[2]};
        r2 = fmaxf(dot3(da, da), fmaxf(dot3(db, db), dot3(dc, dc)));
    }
    r.rad = sqrtf(r2) * 1.0001f + 1e-7f;
    r.pad[0] = r.pad[1] = r.pad[2] = r.pad[3] = 0;
}

// Records are built ONCE per call into stream-ordered scratch (384 B per triangle, L2-resident): building them per
// CTA per tile made two warps of every CTA redo ~500 instructions per triangle while the other six waited at the
// barrier -- half of the kernel's time at 500 k points x 16 k triangles.
__global__ void __launch_bounds__(128)
m2s_records_kernel(const float* __restrict__ tris, const long long num_tris, TriRecord* __restrict__ recs,
                   const int* __restrict__ tri_perm, float4* __restrict__ tsph, unsigned* __restrict__ tflags) {
    const long long t = (long long)blockIdx.x * 128 + threadIdx.x;
    if (t >= num_tris) return;
    TriRecord& r = recs[t];
    build_record(tris + (tri_perm ? (long long)tri_perm[t] : t) * 9, r);
    if (tsph) {                 // compact copy of what the culling levels read: bounding sphere, usable directions
        tsph[t] = make_float4(r.cen[0], r.cen[1], r.cen[2], r.rad);
        tflags[t] = r.dir_ok | (r.nondegenerate ? 0x80000000u : 0u);
    }
}

// Centroids of the triangles, as input of the same counting sort the query points go through: after it a patch of
// M2S_PATCH consecutive records is a small piece of the surface and one bounding sphere describes it well.
__global__ void __launch_bounds__(128)
m2s_centroid_kernel(const float* __restrict__ tris, const long long num_tris, float* __restrict__ cen) {
    const long long t = (long long)blockIdx.x * 128 + threadIdx.x;
    if (t >= num_tris) return;
    const float* v = tris + t * 9;
#pragma unroll
    for (int k = 0; k < 3; ++k) cen[3 * t + k] = (__ldg(v + k) + __ldg(v + 3 + k) + __ldg(v + 6 + k)) * (1.0f / 3.0f);
}

// One bounding sphere per patch: {centre, radius}; the radius is negated when the patch holds no triangle a distance
// can come from (all degenerate), so that it cannot serve as an upper bound.  Culling only.
__global__ void __launch_bounds__(128)
m2s_patch_sphere_kernel(const float4* __restrict__ tsph, const unsigned* __restrict__ tflags, const long long num_tris,
                        float4* __restrict__ spheres, const long long num_patches) {
    const long long patch = (long long)blockIdx.x * 128 + threadIdx.x;
    if (patch >= num_patches) return;
    const long long t0 = patch * M2S_PATCH;
    const int cnt = (int)min((long long)M2S_PATCH, num_tris - t0);
    float c[3] = {0.f, 0.f, 0.f};
    for (int j = 0; j < cnt; ++j) { const float4 s = tsph[t0 + j]; c[0] += s.x; c[1] += s.y; c[2] += s.z; }
    const float inv = 1.0f / (float)cnt;
    c[0] *= inv; c[1] *= inv; c[2] *= inv;
    float rad = 0.f;
    bool any_nondeg = false;
    for (int j = 0; j < cnt; ++j) {
        const float4 s = tsph[t0 + j];
        const float e[3] = {s.x - c[0], s.y - c[1], s.z - c[2]};
        rad = fmaxf(rad, sqrtf(dot3(e, e)) * 1.0001f + s.w);
        any_nondeg |= (tflags[t0 + j] >> 31) != 0;
    }
    rad = rad * 1.0001f + 1e-6f;
    if (!(rad == rad) || !(c[0] == c[0]) || !(c[1] == c[1]) || !(c[2] == c[2])) {       // NaN vertices: never culled
        c[0] = c[1] = c[2] = 0.f; rad = INFINITY; any_nondeg = false;
    }
    spheres[patch] = make_float4(c[0], c[1], c[2], any_nondeg ? rad : -rad);
}

// ---- the exact per-(point, triangle) arithmetic, shared by both kernels ------------------------------------------
// squared distance to the triangle (edge / face branch as in the reference)
__device__ __forceinline__ float m2s_tri_dist2(const TriRecord& r, const float* P, const float* p0) {
    float p1[3], p2[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { p1[k] = P[k] - r.b[k]; p2[k] = P[k] - r.c[k]; }
    const float s1 = copysignf(1.0f, dot3(r.c10, p0));
    const float s2 = copysignf(1.0f, dot3(r.c21, p1));
    const float s3 = copysignf(1.0f, dot3(r.c02, p2));
    float d2;
    if ((s1 + s2 + s3) < 2.0f) {
        const float e1 = d2axmb(r.v10, clamp01(dot3(r.v10, p0) * r.inv10), p0);
        const float e2 = d2axmb(r.v21, clamp01(dot3(r.v21, p1) * r.inv21), p1);
        const float e3 = d2axmb(r.v02, clamp01(dot3(r.v02, p2) * r.inv02), p2);
        d2 = fminf(e1, fminf(e2, e3));
    } else {
        d2 = dot3(r.nor, p0) * dot3(r.nor, p0) * r.invn;
    }
    if (d2 < 0.0f) d2 = 0.0f;
    return d2;
}

// Moller-Trumbore line stabs for the directions in `dirs` (warp-uniform); qvec and edge2.qvec do not depend on the direction
__device__ __forceinline__ void m2s_tri_stabs(const TriRecord& r, const float* p0, const unsigned dirs, unsigned& pos, unsigned& neg) {
    float qvec[3];
    cross3(p0, r.v10, qvec);
    const float edge2[3] = {-r.v02[0], -r.v02[1], -r.v02[2]};
    const float e2q = dot3(edge2, qvec);
#pragma unroll
    for (int k = 0; k < M2S_NDIR; ++k) {
        if (!((dirs >> k) & 1u)) continue;                     // warp-uniform
        const float inv_det = r.inv_det[k];
        const float u = dot3(p0, r.pvec[k]) * inv_det;
        const bool pu = !(u < 0.0f || u > 1.0f);
        if (!__any_sync(0xffffffffu, pu)) continue;            // no lane's line crosses this slab
        const float v = dot3(c_stab_dir[k], qvec) * inv_det;
        const bool pv = pu && !(v < 0.0f || u + v > 1.0f);
        const float t = e2q * inv_det;
        if (pv) { if (t >= 0.0f) pos |= 1u << k; else neg |= 1u << k; }
    }
}

// Bounding sphere of a warp's points (neighbours after the spatial sort).
__device__ __forceinline__ void m2s_warp_sphere(const float* P, const bool active, float* wc, float& wr) {
    const unsigned am = __ballot_sync(0xffffffffu, active);
    const float cnt = (float)max(__popc(am), 1);
    float sx = active ? P[0] : 0.f, sy = active ? P[1] : 0.f, sz = active ? P[2] : 0.f;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o); sz += __shfl_xor_sync(0xffffffffu, sz, o);
    }
    wc[0] = sx / cnt; wc[1] = sy / cnt; wc[2] = sz / cnt;
    const float ex = P[0] - wc[0], ey = P[1] - wc[1], ez = P[2] - wc[2];
    float r2 = active ? ex * ex + ey * ey + ez * ez : 0.f;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) r2 = fmaxf(r2, __shfl_xor_sync(0xffffffffu, r2, o));
    wr = sqrtf(r2) * 1.001f + 1e-6f;
}

__global__ void __launch_bounds__(M2S_THREADS)
mesh2sdf_kernel(const float* __restrict__ points, const long long n, const TriRecord* __restrict__ recs,
                const long long num_tris, float* __restrict__ dist, const int* __restrict__ perm,
                uint3* __restrict__ partial) {
    // gridDim.y > 1: the triangle range is cut into gridDim.y slices (small batches would otherwise leave most SMs
    // waiting on a few CTAs that each walk every triangle); slices merge their running minimum and stab masks through
    // `partial` (atomicMin on the bits of the non-negative squared distance, atomicOr) and m2s_finish_kernel signs them.
    __shared__ TriRecord rec[M2S_TILE];
    static_assert(sizeof(TriRecord) % 16 == 0, "records are copied as 16-byte words");
    const long long i = (long long)blockIdx.x * M2S_THREADS + threadIdx.x;
    const bool active = i < n;
    float P[3] = {0.f, 0.f, 0.f};
    if (active) { P[0] = __ldg(points + 3 * i); P[1] = __ldg(points + 3 * i + 1); P[2] = __ldg(points + 3 * i + 2); }
    float mind2 = INFINITY;
    float mind = INFINITY;              // sqrt(mind2), refreshed when mind2 improves (culling bound only)
    unsigned pos = 0, neg = 0;
    const unsigned all_dirs = (1u << M2S_NDIR) - 1u;
    // Bounding sphere of the warp's points (they are neighbours after the spatial sort): the 13 stab lines of all 32
    // points in direction k lie inside a cylinder of radius wr around the line through wc, so a triangle whose own
    // bounding sphere stays clear of that cylinder cannot be hit by any of them.  Lane k < 13 tests direction k for the
    // whole warp; only the directions that survive run the exact (reference-order) barycentric tests below.  The cull is
    // conservative (generous slack), so every decision and every distance is unchanged.
    const int lane = threadIdx.x & 31;
    float wc[3], wr;
    m2s_warp_sphere(P, active, wc, wr);
    float dk[3] = {0.f, 0.f, 0.f}, inv_dk2 = 0.f;            // this lane's direction (lanes >= 13: none)
    if (lane < M2S_NDIR) {
        dk[0] = c_stab_dir[lane][0]; dk[1] = c_stab_dir[lane][1]; dk[2] = c_stab_dir[lane][2];
        inv_dk2 = 1.0f / dot3(dk, dk);
    }

    const long long per_slice = ((num_tris + gridDim.y - 1) / gridDim.y + M2S_TILE - 1) / M2S_TILE * M2S_TILE;
    const long long t_begin = (long long)blockIdx.y * per_slice;
    const long long t_end = min(num_tris, t_begin + per_slice);
    for (long long t0 = t_begin; t0 < t_end; t0 += M2S_TILE) {
        const int cnt = (int)min((long long)M2S_TILE, t_end - t0);
        __syncthreads();
        {
            const uint4* src = reinterpret_cast<const uint4*>(recs + t0);
            uint4* dst = reinterpret_cast<uint4*>(rec);
            for (int e = threadIdx.x; e < cnt * (int)(sizeof(TriRecord) / 16); e += M2S_THREADS) dst[e] = __ldg(src + e);
        }
        __syncthreads();
        // All 32 lanes stay in the loop (inactive lanes carry P = 0 and never write): the rejections below are
        // warp-votes, so the branches are uniform and most (triangle, direction) pairs cost 5 instructions.
#pragma unroll (M2S_UNROLL)
        for (int tt = 0; tt < cnt; ++tt) {
            const TriRecord& r = rec[tt];
            float p0[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) p0[k] = P[k] - r.a[k];
            // Exact-preserving cull: every point of the triangle is at least |P - cen| - rad away, so if that already
            // exceeds the running minimum the edge/face distance (the expensive half) cannot lower it.
            const float pc[3] = {P[0] - r.cen[0], P[1] - r.cen[1], P[2] - r.cen[2]};
            const float reach = mind + r.rad;
            const bool near = r.nondegenerate && !(dot3(pc, pc) > reach * reach * 1.00001f);
            if (__any_sync(0xffffffffu, near)) {
                const float d2 = m2s_tri_dist2(r, P, p0);
                if (near && d2 < mind2) { mind2 = d2; mind = sqrtf(d2) * 1.00001f; }
            }
            // 13 line stabs (Moller-Trumbore); qvec and edge2.qvec do not depend on the direction
            const bool undecided = (pos & neg) != all_dirs;
            if (r.dir_ok && __any_sync(0xffffffffu, undecided)) {
                // warp-cylinder cull, one direction per lane
                bool maybe = false;
                if (lane < M2S_NDIR) {
                    const float vx = r.cen[0] - wc[0], vy = r.cen[1] - wc[1], vz = r.cen[2] - wc[2];
                    const float proj = vx * dk[0] + vy * dk[1] + vz * dk[2];
                    const float perp2 = (vx * vx + vy * vy + vz * vz) - proj * proj * inv_dk2;
                    const float reach = (r.rad + wr) * 1.001f + 1e-5f;
                    maybe = !(perp2 > reach * reach);
                }
                const unsigned dirs = __ballot_sync(0xffffffffu, maybe) & r.dir_ok;
                if (!dirs) continue;
                m2s_tri_stabs(r, p0, dirs, pos, neg);
            }
        }
    }
    if (active) {
        if (mind2 < 0.0f) mind2 = 0.0f;
        if (partial) {
            atomicMin(&partial[i].x, __float_as_uint(mind2));
            if (pos) atomicOr(&partial[i].y, pos);
            if (neg) atomicOr(&partial[i].z, neg);
            return;
        }
        float d = sqrtf(mind2);
        if ((pos & neg) == all_dirs) d = -d;
        dist[perm ? (long long)perm[i] : i] = d;
    }
}

// ---- large batches: distance and sign by two output-sensitive kernels ---------------------------------------------
// The brute-force walk above spends almost all of its time dismissing (warp, triangle) pairs one triangle at a time.
// For large batches the two halves of the answer are computed separately, each touching only the pairs that matter;
// the exact per-pair arithmetic is the same code, and min / OR do not depend on which dismissed pairs were skipped or on
// the order, so every distance and every sign equals the brute-force walk bit for bit (A/B test in tests/).
//
// (1) DISTANCE, point-driven.  The records are sorted along a Morton curve (a patch of 32 consecutive records is a small
//     piece of surface).  A warp, on its own (no shared memory, no barriers): one PATCH per lane -- keep it if its sphere
//     could hold a closer triangle for one of the warp's points; for every kept patch one TRIANGLE per lane, the same test
//     on the triangle's sphere; the survivors, one after the other, through the exact arithmetic.  The bound starts from
//     a pre-pass (each point's exact distance to the triangles of the two patches nearest to its warp) and tightens as the
//     warp goes; the patch range is cut into slices over gridDim.y and the slices merge by atomicMin.
//
// (2) SIGN, triangle-driven.  A line through P along stab direction k hits a triangle only if P's projection along k
//     falls inside the triangle's projection.  For each of the 13 directions the points are binned on a G x G grid of the
//     plane across it, spanning the extent of the call's points and vertices (one counting sort over all 13 x G x G
//     cells; exactly 13 n entries, so the scratch is bounded without a host round trip).  One warp per (triangle, direction) walks the cells under the triangle's projected
//     bounding box -- widened by the worst rounding error of the exact test, which grows as the line grazes the
//     triangle's plane -- and runs the exact test on the points binned there, one point per lane, OR-ing the 13-bit
//     masks of the points it hits.
constexpr int M2S_WARP_THREADS = 128;
#ifndef NGLOD_M2S_HIER_MIN
#define NGLOD_M2S_HIER_MIN 16384
#endif
constexpr long long M2S_HIER_MIN_POINTS = NGLOD_M2S_HIER_MIN;
constexpr float M2S_UNIT_R = 1.8f;          // scene radius the rounding-error constants below were derived for (unit cube: sqrt(3))

// The projected grids span [-R, R]^2 with R = the largest |point| or |vertex| of the call (a device-side value: no host
// round trip), so a mesh that was not normalised into the unit sphere still spreads over the cells.
__global__ void __launch_bounds__(256)
m2s_extent_kernel(const float* __restrict__ points, const long long n, const float* __restrict__ tris, const long long num_tris,
                  unsigned* __restrict__ ext2_bits) {
    const long long total = n + 3 * num_tris;
    float m = 0.f;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const float* v = i < n ? points + 3 * i : tris + 3 * (i - n);
        const float r2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
        if (r2 == r2) m = fmaxf(m, r2);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(ext2_bits, __float_as_uint(m));      // non-negative floats order like their bits
}
__device__ __forceinline__ float m2s_proj_radius(const unsigned* __restrict__ ext2_bits) {
    return fmaxf(sqrtf(__uint_as_float(__ldg(ext2_bits))) * 1.001f, 1e-3f);
}

// Pre-pass of the distance half: every point's exact distance to the triangles of the two patches nearest to its warp.
// That is a true candidate of the minimum (it goes into partial[].x like any other) and a tight culling bound for
// every slice of the walk below from its first patch on -- with the bound read off the patch spheres alone, a slice
// that does not contain the near part of the surface kept every triangle within (distance + patch radius + 2 warp
// radii) alive for the whole slice.
__global__ void __launch_bounds__(M2S_WARP_THREADS)
mesh2sdf_bound_kernel(const float* __restrict__ points, const long long n, const TriRecord* __restrict__ recs,
                      const unsigned* __restrict__ tflags, const float4* __restrict__ patches,
                      const long long num_tris, uint3* __restrict__ partial) {
    const long long i = (long long)blockIdx.x * M2S_WARP_THREADS + threadIdx.x;
    const bool active = i < n;
    const int lane = threadIdx.x & 31;
    float P[3] = {0.f, 0.f, 0.f};
    if (active) { P[0] = __ldg(points + 3 * i); P[1] = __ldg(points + 3 * i + 1); P[2] = __ldg(points + 3 * i + 2); }
    {
        const unsigned am = __ballot_sync(0xffffffffu, active);
        if (am == 0u) return;
        const int src = __ffs(am) - 1;
        const float q0 = __shfl_sync(0xffffffffu, P[0], src), q1 = __shfl_sync(0xffffffffu, P[1], src), q2 = __shfl_sync(0xffffffffu, P[2], src);
        if (!active) { P[0] = q0; P[1] = q1; P[2] = q2; }
    }
    float wc[3], wr;
    m2s_warp_sphere(P, true, wc, wr);
    const long long num_patches = (num_tris + M2S_PATCH - 1) / M2S_PATCH;
    // the two patches whose spheres come closest to the warp's centre (only patches a distance can come from)
    float k1 = INFINITY, k2 = INFINITY;
    long long p1 = -1, p2 = -1;
    for (long long pt = lane; pt < num_patches; pt += 32) {
        const float4 s = __ldg(patches + pt);
        if (!(s.w >= 0.0f)) continue;
        const float v[3] = {s.x - wc[0], s.y - wc[1], s.z - wc[2]};
        const float key = sqrtf(dot3(v, v)) - s.w;
        if (key < k1) { k2 = k1; p2 = p1; k1 = key; p1 = pt; }
        else if (key < k2) { k2 = key; p2 = pt; }
    }
    long long pick[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        float bk = k1; long long bp = p1;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const float ok = __shfl_xor_sync(0xffffffffu, bk, o);
            const long long op = __shfl_xor_sync(0xffffffffu, bp, o);
            if (ok < bk || (ok == bk && op >= 0 && (bp < 0 || op < bp))) { bk = ok; bp = op; }
        }
        pick[r] = bp;
        if (bp >= 0 && bp == p1) { k1 = k2; p1 = p2; k2 = INFINITY; p2 = -1; }      // the owner moves on to its runner-up
    }
    float mind2 = INFINITY;
#pragma unroll 1
    for (int r = 0; r < 2; ++r) {
        if (pick[r] < 0) continue;
        const long long t0 = pick[r] * M2S_PATCH;
        const unsigned fl = (t0 + lane < num_tris) ? __ldg(tflags + t0 + lane) : 0u;
        unsigned todo = __ballot_sync(0xffffffffu, (fl >> 31) != 0u);
        while (todo) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            const TriRecord& rr = recs[t0 + j];
            float p0[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) p0[k] = P[k] - rr.a[k];
            mind2 = fminf(mind2, m2s_tri_dist2(rr, P, p0));
        }
    }
    if (active && mind2 < INFINITY) partial[i].x = __float_as_uint(mind2);       // after m2s_init_kernel, before the walk
}

__global__ void __launch_bounds__(M2S_WARP_THREADS)
mesh2sdf_dist_kernel(const float* __restrict__ points, const long long n, const TriRecord* __restrict__ recs,
                     const float4* __restrict__ tsph, const unsigned* __restrict__ tflags,
                     const float4* __restrict__ patches, const long long num_tris, uint3* __restrict__ partial) {
    const long long i = (long long)blockIdx.x * M2S_WARP_THREADS + threadIdx.x;
    const bool active = i < n;
    const int lane = threadIdx.x & 31;
    float P[3] = {0.f, 0.f, 0.f};
    if (active) { P[0] = __ldg(points + 3 * i); P[1] = __ldg(points + 3 * i + 1); P[2] = __ldg(points + 3 * i + 2); }
    {   // lanes past the end shadow the warp's first point: they never write and must not widen the votes below
        const unsigned am = __ballot_sync(0xffffffffu, active);
        if (am == 0u) return;
        const int src = __ffs(am) - 1;
        const float q0 = __shfl_sync(0xffffffffu, P[0], src), q1 = __shfl_sync(0xffffffffu, P[1], src), q2 = __shfl_sync(0xffffffffu, P[2], src);
        if (!active) { P[0] = q0; P[1] = q1; P[2] = q2; }
    }
    float wc[3], wr;
    m2s_warp_sphere(P, true, wc, wr);
    float mind2 = INFINITY;
    const long long num_patches = (num_tris + M2S_PATCH - 1) / M2S_PATCH;
    // culling bound: the lane's distance to the triangles the pre-pass looked at (a true candidate, hence an upper bound)
    float mind;
    {
        const unsigned am = __ballot_sync(0xffffffffu, active);
        float m0 = active ? __uint_as_float(partial[i].x) : 0.f;
        const float ms = __shfl_sync(0xffffffffu, m0, __ffs(am) - 1);
        if (!active) m0 = ms;
        mind = sqrtf(m0) * 1.00001f + 1e-7f;
    }
    // gridDim.y slices of the patch range: points near the medial axis (the centre of a sphere, the axis of a torus) are
    // equally far from a large part of the surface, no bound can dismiss it, and the few warps holding them would walk
    // tens of thousands of triangles while the machine idles; sliced, that walk is spread over gridDim.y warps.
    const long long per_slice = (num_patches + gridDim.y - 1) / gridDim.y;
    const long long patch_begin = (long long)blockIdx.y * per_slice;
    const long long patch_end = min(num_patches, patch_begin + per_slice);
    float wmind = mind;             // loosest running bound in the warp
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) wmind = fmaxf(wmind, __shfl_xor_sync(0xffffffffu, wmind, o));
    for (long long pbase = patch_begin; pbase < patch_end; pbase += 32) {
        bool keep = false;
        if (pbase + lane < patch_end) {
            const float4 s = __ldg(patches + pbase + lane);
            const float v[3] = {s.x - wc[0], s.y - wc[1], s.z - wc[2]};
            const float reach = (wmind + fabsf(s.w) + wr) * 1.001f + 1e-5f;
            keep = !(dot3(v, v) > reach * reach);
        }
        unsigned pmask = __ballot_sync(0xffffffffu, keep);
        while (pmask) {
            const int pb = __ffs(pmask) - 1;
            pmask &= pmask - 1;
            const long long t0 = (pbase + pb) * M2S_PATCH;
            bool near_l = false;
            if (t0 + lane < num_tris && (__ldg(tflags + t0 + lane) >> 31)) {
                const float4 s = __ldg(tsph + t0 + lane);
                const float v[3] = {s.x - wc[0], s.y - wc[1], s.z - wc[2]};
                const float reach = (wmind + s.w + wr) * 1.001f + 1e-5f;
                near_l = !(dot3(v, v) > reach * reach);
            }
            unsigned surv = __ballot_sync(0xffffffffu, near_l);
            if (!surv) continue;
            while (surv) {
                const int j = __ffs(surv) - 1;
                surv &= surv - 1;
                const TriRecord& r = recs[t0 + j];
                const float pc[3] = {P[0] - r.cen[0], P[1] - r.cen[1], P[2] - r.cen[2]};
                const float reach = mind + r.rad;
                const bool near = !(dot3(pc, pc) > reach * reach * 1.00001f);
                if (__any_sync(0xffffffffu, near)) {
                    float p0[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) p0[k] = P[k] - r.a[k];
                    const float d2 = m2s_tri_dist2(r, P, p0);
                    if (near && d2 < mind2) { mind2 = d2; mind = fminf(mind, sqrtf(d2) * 1.00001f); }
                }
            }
            wmind = mind;
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) wmind = fmaxf(wmind, __shfl_xor_sync(0xffffffffu, wmind, o));
        }
    }
    if (active && mind2 < INFINITY) atomicMin(&partial[i].x, __float_as_uint(mind2));       // .y / .z belong to the stab kernel
}

// two unit vectors across stab direction k (any fixed pair works: points and triangles use the same one)
__device__ __forceinline__ void m2s_stab_basis(const int k, float* e1, float* e2) {
    const float d[3] = {c_stab_dir[k][0], c_stab_dir[k][1], c_stab_dir[k][2]};
    const float ax = fabsf(d[0]), ay = fabsf(d[1]), az = fabsf(d[2]);
    float a[3] = {0.f, 0.f, 0.f};
    if (ax <= ay && ax <= az) a[0] = 1.f; else if (ay <= az) a[1] = 1.f; else a[2] = 1.f;
    cross3(d, a, e1);
    const float n1 = 1.0f / sqrtf(dot3(e1, e1));
    e1[0] *= n1; e1[1] *= n1; e1[2] *= n1;
    cross3(d, e1, e2);
    const float n2 = 1.0f / sqrtf(dot3(e2, e2));
    e2[0] *= n2; e2[1] *= n2; e2[2] *= n2;
}

// cell coordinate of a projected coordinate: monotone in c, clamped to the grid (NaN -> 0)
__device__ __forceinline__ int m2s_proj_cell(const float c, const float R, const float inv_h, const int G) {
    const float f = floorf((c + R) * inv_h);
    return (int)fminf(fmaxf(f, 0.0f), (float)(G - 1));
}

// pass 1 (cursor == counts, pidx == nullptr): histogram of the 13 projected cells of every point;
// pass 2 (cursor == exclusive offsets): scatter the point indices; afterwards cursor[c] is the END of cell c.
__global__ void __launch_bounds__(256)
m2s_proj_bin_kernel(const float* __restrict__ x, const long long n, const int G, const unsigned* __restrict__ ext2_bits,
                    int* __restrict__ cursor, int* __restrict__ pidx) {
    __shared__ float basis[M2S_NDIR][6];
    if (threadIdx.x < M2S_NDIR) m2s_stab_basis(threadIdx.x, basis[threadIdx.x], basis[threadIdx.x] + 3);
    __syncthreads();
    const float R = m2s_proj_radius(ext2_bits);
    const float inv_h = (float)G / (2.0f * R);
    const long long stride = (long long)gridDim.x * 256;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
        const float P[3] = {__ldg(x + 3 * i), __ldg(x + 3 * i + 1), __ldg(x + 3 * i + 2)};
#pragma unroll 1
        for (int k = 0; k < M2S_NDIR; ++k) {
            const int cx = m2s_proj_cell(dot3(P, basis[k]), R, inv_h, G);
            const int cy = m2s_proj_cell(dot3(P, basis[k] + 3), R, inv_h, G);
            const int slot = atomicAdd(cursor + ((long long)k * G + cy) * G + cx, 1);
            if (pidx) pidx[slot] = (int)i;
        }
    }
}


__global__ void __launch_bounds__(M2S_WARP_THREADS)
mesh2sdf_stab_kernel(const float* __restrict__ points, const TriRecord* __restrict__ recs, const long long num_tris,
                     const int G, const unsigned* __restrict__ ext2_bits, const int* __restrict__ cell_end,
                     const int* __restrict__ pidx, uint3* __restrict__ partial
                     ) {
    const int lane = threadIdx.x & 31;
    const long long task = (long long)blockIdx.x * (M2S_WARP_THREADS / 32) + (threadIdx.x >> 5);
    const long long t = task / M2S_NDIR;
    const int k = (int)(task - t * M2S_NDIR);
    if (t >= num_tris) return;
    const TriRecord& r = recs[t];
    if (!((r.dir_ok >> k) & 1u)) return;
    float e1[3], e2[3];
    m2s_stab_basis(k, e1, e2);
    const float inv_det = r.inv_det[k];
    // widest rounding error of u / v, as a displacement along the edges: ~eps |p0| |v10| |v02| / |det|; the constants are
    // for |p0| <= ~4 (unit cube) and scale with the radius of the scene
    const float R = m2s_proj_radius(ext2_bits);
    const float scale = fmaxf(1.0f, R * (1.0f / M2S_UNIT_R));
    float m = (1e-5f + 2e-6f * sqrtf(dot3(r.v10, r.v10) * dot3(r.v02, r.v02)) * fabsf(inv_det)) * 1.01f * scale;
    if (!(m < 4.0f * R)) m = 4.0f * R;                           // also catches NaN: walk the whole grid
    const float xa = dot3(r.a, e1), xb = dot3(r.b, e1), xc = dot3(r.c, e1);
    const float ya = dot3(r.a, e2), yb = dot3(r.b, e2), yc = dot3(r.c, e2);
    const float inv_h = (float)G / (2.0f * R);
    int cx0 = m2s_proj_cell(fminf(xa, fminf(xb, xc)) - m, R, inv_h, G), cx1 = m2s_proj_cell(fmaxf(xa, fmaxf(xb, xc)) + m, R, inv_h, G);
    int cy0 = m2s_proj_cell(fminf(ya, fminf(yb, yc)) - m, R, inv_h, G), cy1 = m2s_proj_cell(fmaxf(ya, fmaxf(yb, yc)) + m, R, inv_h, G);
    if (!(xa == xa && xb == xb && xc == xc && ya == ya && yb == yb && yc == yc)) { cx0 = cy0 = 0; cx1 = cy1 = G - 1; }
    const float edge2[3] = {-r.v02[0], -r.v02[1], -r.v02[2]};
    const float pv[3] = {r.pvec[k][0], r.pvec[k][1], r.pvec[k][2]};
    const float a[3] = {r.a[0], r.a[1], r.a[2]};
    const float v10[3] = {r.v10[0], r.v10[1], r.v10[2]};
    const unsigned bit = 1u << k;
    for (int cy = cy0; cy <= cy1; ++cy) {
        const long long c0 = ((long long)k * G + cy) * G + cx0, c1 = ((long long)k * G + cy) * G + cx1;
        const int begin = c0 ? __ldg(cell_end + c0 - 1) : 0;
        const int end = __ldg(cell_end + c1);
        for (int jj = begin + lane; jj < end; jj += 32) {
            const long long i = __ldg(pidx + jj);
            const float P[3] = {__ldg(points + 3 * i), __ldg(points + 3 * i + 1), __ldg(points + 3 * i + 2)};
            float p0[3];
#pragma unroll
            for (int q = 0; q < 3; ++q) p0[q] = P[q] - a[q];
            // the same expressions as m2s_tri_stabs
            const float u = dot3(p0, pv) * inv_det;
            if (u < 0.0f || u > 1.0f) continue;
            float qvec[3];
            cross3(p0, v10, qvec);
            const float e2q = dot3(edge2, qvec);
            const float v = dot3(c_stab_dir[k], qvec) * inv_det;
            if (v < 0.0f || u + v > 1.0f) continue;
            const float tt = e2q * inv_det;
            if (tt >= 0.0f) atomicOr(&partial[i].y, bit); else atomicOr(&partial[i].z, bit);
        }
    }
}

__global__ void __launch_bounds__(256)
m2s_init_kernel(uint3* __restrict__ partial, const long long n) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i < n) partial[i] = make_uint3(0x7f800000u, 0u, 0u);       // {+inf, no stab hits}
}

__global__ void __launch_bounds__(256)
m2s_finish_kernel(const uint3* __restrict__ partial, const long long n, float* __restrict__ dist, const int* __restrict__ perm) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const uint3 p = partial[i];
    float d = sqrtf(__uint_as_float(p.x));
    if ((p.y & p.z) == (1u << M2S_NDIR) - 1u) d = -d;
    dist[perm ? (long long)perm[i] : i] = d;
}

// ---- spatial counting sort of the query points (32^3 Morton bins) ------------------------------------------------
// The rejections above are warp votes, so they only pay off when the 32 points of a warp are close together: then
// almost every (triangle, warp) pair is dismissed by the bounding-sphere cull and almost every stab direction by the
// first barycentric test.  The sampler hands points over in random spatial order, so they are grouped first
// (histogram -> scan -> scatter on stream-ordered scratch); distances are written back through the permutation and
// every value is unchanged.
#ifndef NGLOD_M2S_BIN_BITS
#define NGLOD_M2S_BIN_BITS 5
#endif
constexpr int M2S_BIN_BITS = NGLOD_M2S_BIN_BITS;
constexpr int M2S_BIN_RES = 1 << M2S_BIN_BITS;
constexpr int M2S_BINS = M2S_BIN_RES * M2S_BIN_RES * M2S_BIN_RES;

__device__ __forceinline__ int m2s_bin(float x, float y, float z) {
    const int bx = min(M2S_BIN_RES - 1, max(0, (int)floorf((x + 1.f) * (0.5f * M2S_BIN_RES))));
    const int by = min(M2S_BIN_RES - 1, max(0, (int)floorf((y + 1.f) * (0.5f * M2S_BIN_RES))));
    const int bz = min(M2S_BIN_RES - 1, max(0, (int)floorf((z + 1.f) * (0.5f * M2S_BIN_RES))));
    int code = 0;
#pragma unroll
    for (int b = 0; b < M2S_BIN_BITS; ++b) code |= (((bx >> b) & 1) << (3 * b + 2)) | (((by >> b) & 1) << (3 * b + 1)) | (((bz >> b) & 1) << (3 * b));
    return code;
}

__global__ void __launch_bounds__(256)
m2s_hist_kernel(const float* __restrict__ x, const long long n, int* __restrict__ hist) {
    const long long stride = (long long)gridDim.x * 256;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += stride)
        atomicAdd(hist + m2s_bin(__ldg(x + 3 * i), __ldg(x + 3 * i + 1), __ldg(x + 3 * i + 2)), 1);
}

__global__ void __launch_bounds__(1024)
m2s_scan_kernel(int* __restrict__ hist, const int nbins) {      // in place: counts -> exclusive offsets (nbins % 1024 == 0, one block)
    __shared__ int warp_sums[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < nbins; base += 1024) {
        const int v = hist[base + threadIdx.x];
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const int tot = incl + (warp ? warp_sums[warp - 1] : 0) + carry;
        hist[base + threadIdx.x] = tot - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = tot;
        __syncthreads();
    }
}

// Large bin counts (the 13 projected grids): 1024 bins per block, then the block totals through the one-block scan above.
__global__ void __launch_bounds__(1024)
m2s_scan_local_kernel(int* __restrict__ hist, int* __restrict__ block_sums) {       // gridDim.x * 1024 bins
    __shared__ int warp_sums[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long idx = (long long)blockIdx.x * 1024 + threadIdx.x;
    const int v = hist[idx];
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = warp_sums[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
        warp_sums[lane] = w;
    }
    __syncthreads();
    const int tot = incl + (warp ? warp_sums[warp - 1] : 0);
    hist[idx] = tot - v;
    if (threadIdx.x == 1023) block_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024)
m2s_scan_add_kernel(int* __restrict__ hist, const int* __restrict__ block_offsets) {
    hist[(long long)blockIdx.x * 1024 + threadIdx.x] += block_offsets[blockIdx.x];
}

__global__ void __launch_bounds__(256)
m2s_scatter_kernel(const float* __restrict__ x, const long long n, int* __restrict__ cursor, float* __restrict__ xs,
                   int* __restrict__ perm) {
    const long long stride = (long long)gridDim.x * 256;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
        const float px = __ldg(x + 3 * i), py = __ldg(x + 3 * i + 1), pz = __ldg(x + 3 * i + 2);
        const int pos = atomicAdd(cursor + m2s_bin(px, py, pz), 1);
        xs[3 * (long long)pos] = px; xs[3 * (long long)pos + 1] = py; xs[3 * (long long)pos + 2] = pz;
        perm[pos] = (int)i;
    }
}

// Adam on a flat buffer (torch.optim.Adam semantics, no amsgrad / weight decay).  HBM-bound: 28 bytes per parameter (reads
// p, g, m, v, writes p, m, v), moved as 16-byte vectors, 4 independent loads in flight per operand and thread.
struct AdamK { float b1, b2, eps, bc2_sqrt, step_size; };
__device__ __forceinline__ void adam_one(float& p, const float g, float& m, float& v, const AdamK& k) {
    const float mi = m + (g - m) * (1.f - k.b1);                       // lerp, as torch's single-tensor Adam
    const float vi = v * k.b2 + g * g * (1.f - k.b2);
    m = mi; v = vi;
    const float denom = sqrtf(vi) / k.bc2_sqrt + k.eps;
    p = p - k.step_size * (mi / denom);
}
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            const long long n, const float lr, const float b1, const float b2, const float eps,
            const float bc1, const float bc2_sqrt) {
    const AdamK k = {b1, b2, eps, bc2_sqrt, lr / bc1};
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // 16-byte vectors when all four buffers are 16-byte aligned (torch allocations and 4-element-aligned slices of them);
    // otherwise, and for the last n % 4 elements, the scalar loop
    const bool aligned = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                           reinterpret_cast<uintptr_t>(v)) & 15u) == 0;
    const long long n4 = aligned ? n >> 2 : 0;
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    for (long long i = tid; i < n4; i += stride) {
        float4 pp = p4[i], mm = m4[i], vv = v4[i];
        const float4 gg = __ldcs(g4 + i);                            // gradients are dead after this step
        adam_one(pp.x, gg.x, mm.x, vv.x, k); adam_one(pp.y, gg.y, mm.y, vv.y, k);
        adam_one(pp.z, gg.z, mm.z, vv.z, k); adam_one(pp.w, gg.w, mm.w, vv.w, k);
        p4[i] = pp; m4[i] = mm; v4[i] = vv;
    }
    for (long long i = (n4 << 2) + tid; i < n; i += stride) {
        float pi = p[i], mi = m[i], vi = v[i];
        adam_one(pi, g[i], mi, vi, k);
        p[i] = pi; m[i] = mi; v[i] = vi;
    }
}

}  // namespace

// Stream-ordered scratch comes from a pool of the library's own, one per device, that keeps what it is handed back: the
// device's default pool trims itself at every synchronisation, and a resample that re-allocates its ~50-150 MB of
// records / bins after each trim cost 3 ms on top of 2.5 ms of kernels.
static int m2s_scratch_pool(cudaMemPool_t* out) {
    static cudaMemPool_t pools[64] = {};
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    int dev = 0;
    NGLOD_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return NGLOD_EINVAL;
    if (!pools[dev]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        cudaMemPool_t p = nullptr;
        NGLOD_CUDA_TRY(cudaMemPoolCreate(&p, &props));
        unsigned long long keep = ~0ull;
        NGLOD_CUDA_TRY(cudaMemPoolSetAttribute(p, cudaMemPoolAttrReleaseThreshold, &keep));
        pools[dev] = p;
    }
    *out = pools[dev];
    return 0;
}

// exclusive scan of nbins counts in place (nbins % 1024 == 0): per 1024-bin block, the block totals, add back
static void m2s_scan(int* bins, long long nbins, int* bsum, int bsum_n, cudaStream_t st) {
    const int nblk = (int)(nbins / 1024);
    cudaMemsetAsync(bsum, 0, (size_t)bsum_n * 4, st);
    m2s_scan_local_kernel<<<nblk, 1024, 0, st>>>(bins, bsum);
    m2s_scan_kernel<<<1, 1024, 0, st>>>(bsum, bsum_n);
    m2s_scan_add_kernel<<<nblk, 1024, 0, st>>>(bins, bsum);
}

extern "C" int nglod_release_scratch(void) {
    cudaMemPool_t pool = nullptr;
    if (int e = m2s_scratch_pool(&pool)) return e;
    return (int)cudaMemPoolTrimTo(pool, 0);
}

extern "C" int nglod_mesh2sdf(const float* points, int64_t n, const float* tris, int64_t num_tris, float* dist,
                              void* stream) {
    return nglod_mesh2sdf_ex(points, n, tris, num_tris, dist, 0u, stream);
}

extern "C" int nglod_mesh2sdf_ex(const float* points, int64_t n, const float* tris, int64_t num_tris, float* dist,
                                 uint32_t flags, void* stream) {
    if (n < 0 || num_tris < 0 || (n > 0 && (!points || !dist)) || (num_tris > 0 && !tris)) return NGLOD_EINVAL;
    if (n == 0) return 0;
    const long long grid = (n + M2S_THREADS - 1) / M2S_THREADS;
    if (grid > 2147483647ll) return NGLOD_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (num_tris == 0) {            // no surface: the reference's min over nothing -- keep the old kernel's answer (+inf)
        mesh2sdf_kernel<<<(int)grid, M2S_THREADS, 0, st>>>(points, (long long)n, nullptr, 0ll, dist, nullptr, nullptr);
        return (int)cudaGetLastError();
    }
    const bool sort = !(n < 4096 || n >= 2000000000ll || num_tris < 64);      // else too small for the sort to pay
    // Large batches: distance and sign by the two output-sensitive kernels.  Their set-up (two more counting sorts, the
    // 13 projected grids) costs a fixed ~0.5 ms, so small batches keep the sliced brute-force walk.
    // NGLOD_M2S_FORCE_WALK forces the walk: the A/B switch of the parity test.
    const bool hier = !(flags & NGLOD_M2S_FORCE_WALK) && sort && n >= M2S_HIER_MIN_POINTS &&
                      num_tris >= 4 * M2S_PATCH && num_tris < 100000000ll && n * M2S_NDIR < 2000000000ll;
    const long long num_patches = (num_tris + M2S_PATCH - 1) / M2S_PATCH;
    int G = 32;                                                              // projected grid: ~2 sqrt(#triangles) cells a side
    while (G < 512 && (long long)G * G < 4 * num_tris) G *= 2;               // sweep x1 / x4 / x16 / x64 in profiles/README.md
    const long long proj_bins = (long long)M2S_NDIR * G * G;                 // a multiple of 1024
    // slices of the triangle range (walk only): enough CTAs for ~4 waves of the machine, at least 4 tiles per slice
    int slices = 1;
    if (!hier) {
        const long long ctas_per_wave = (long long)nglod_sm_count() * 5;
        long long want = (4 * ctas_per_wave + grid - 1) / grid;
        const long long max_slices = num_tris / (4 * M2S_TILE);
        if (want > max_slices) want = max_slices;
        if (want > 32) want = 32;
        if (want > 1) slices = (int)want;
    }
    size_t ws_bytes = 0;
    auto reserve = [&ws_bytes](size_t bytes) { const size_t off = ws_bytes; ws_bytes += (bytes + 255) & ~(size_t)255; return off; };
    const size_t rec_off = reserve((size_t)num_tris * sizeof(TriRecord));
    const size_t part_off = reserve(slices > 1 || hier ? (size_t)n * sizeof(uint3) : 0);
    const size_t xs_off = reserve(sort ? (size_t)n * 12 : 0);
    const size_t perm_off = reserve(sort ? (size_t)n * 4 : 0);
    const size_t hist_off = reserve(sort ? (size_t)M2S_BINS * 4 : 0);
    const size_t tcen_off = reserve(hier ? (size_t)num_tris * 12 : 0);
    const size_t tcs_off = reserve(hier ? (size_t)num_tris * 12 : 0);
    const size_t tperm_off = reserve(hier ? (size_t)num_tris * 4 : 0);
    const size_t tsph_off = reserve(hier ? (size_t)num_tris * sizeof(float4) : 0);
    const size_t tfl_off = reserve(hier ? (size_t)num_tris * 4 : 0);
    const size_t psph_off = reserve(hier ? (size_t)num_patches * sizeof(float4) : 0);
    const size_t pbin_off = reserve(hier ? (size_t)proj_bins * 4 : 0);
    const size_t pidx_off = reserve(hier ? (size_t)n * M2S_NDIR * 4 : 0);
    const size_t ext_off = reserve(hier ? 4 : 0);
    const int bsum_n = (int)((std::max<long long>(proj_bins, M2S_BINS) / 1024 + 1023) / 1024 * 1024);
    const size_t bsum_off = reserve((size_t)bsum_n * 4);
    // slices of the patch range for the distance kernel.  500 k points: 4 / 8 / 16 / 32 / 64 slices measured 3.3 / 2.7 / 2.5 /
    // 2.6 / 3.0 ms (profiles/README.md); a data-parallel rank's 62.5 k points: 16 / 32 / 64 / 128 slices 0.87 / 0.74 / 0.67 /
    // 0.67 ms -- a small batch has too few warps to hide the long walks of its medial-axis points, so it is cut finer
    long long dist_slices = num_patches / 8;                                  // >= 8 patches per slice
    const long long slice_cap = n < 200000 ? 64 : 32;
    if (dist_slices > slice_cap) dist_slices = slice_cap;
    if (dist_slices < 1) dist_slices = 1;
    char* ws = nullptr;
    cudaMemPool_t pool = nullptr;
    if (int e = m2s_scratch_pool(&pool)) return e;
    NGLOD_CUDA_TRY(cudaMallocFromPoolAsync(&ws, ws_bytes + 256, pool, st));
    TriRecord* recs = reinterpret_cast<TriRecord*>(ws + rec_off);
    uint3* partial = (slices > 1 || hier) ? reinterpret_cast<uint3*>(ws + part_off) : nullptr;
    int* hist = reinterpret_cast<int*>(ws + hist_off);
    const long long sm8 = (long long)nglod_sm_count() * 8;
    int err = 0;
    const int* tri_perm = nullptr;
    float4* tsph = hier ? reinterpret_cast<float4*>(ws + tsph_off) : nullptr;
    unsigned* tflags = hier ? reinterpret_cast<unsigned*>(ws + tfl_off) : nullptr;
    float4* patches = hier ? reinterpret_cast<float4*>(ws + psph_off) : nullptr;
    if (hier) {                     // triangles along a Morton curve (the same counting sort, on their centroids)
        float* tcen = reinterpret_cast<float*>(ws + tcen_off);
        int* tp = reinterpret_cast<int*>(ws + tperm_off);
        err = (int)cudaMemsetAsync(hist, 0, (size_t)M2S_BINS * 4, st);
        if (!err) {
            const int nb = (int)std::min<long long>((num_tris + 255) / 256, sm8);
            m2s_centroid_kernel<<<(int)((num_tris + 127) / 128), 128, 0, st>>>(tris, (long long)num_tris, tcen);
            m2s_hist_kernel<<<nb, 256, 0, st>>>(tcen, (long long)num_tris, hist);
            m2s_scan(hist, M2S_BINS, reinterpret_cast<int*>(ws + bsum_off), bsum_n, st);
            m2s_scatter_kernel<<<nb, 256, 0, st>>>(tcen, (long long)num_tris, hist, reinterpret_cast<float*>(ws + tcs_off), tp);
            err = (int)cudaGetLastError();
            tri_perm = tp;
        }
    }
    if (!err) {
        m2s_records_kernel<<<(int)((num_tris + 127) / 128), 128, 0, st>>>(tris, (long long)num_tris, recs, tri_perm, tsph, tflags);
        err = (int)cudaGetLastError();
    }
    if (!err && hier) {
        m2s_patch_sphere_kernel<<<(int)((num_patches + 127) / 128), 128, 0, st>>>(tsph, tflags, (long long)num_tris, patches, num_patches);
        err = (int)cudaGetLastError();
    }
    if (!err && partial) {
        m2s_init_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(partial, (long long)n);
        err = (int)cudaGetLastError();
    }
    const float* pts = points;
    const int* perm = nullptr;
    if (!err && sort) {
        float* xs = reinterpret_cast<float*>(ws + xs_off);
        int* pm = reinterpret_cast<int*>(ws + perm_off);
        err = (int)cudaMemsetAsync(hist, 0, (size_t)M2S_BINS * 4, st);
        if (!err) {
            const int nb = (int)std::min<long long>((n + 255) / 256, sm8);
            m2s_hist_kernel<<<nb, 256, 0, st>>>(points, (long long)n, hist);
            m2s_scan(hist, M2S_BINS, reinterpret_cast<int*>(ws + bsum_off), bsum_n, st);
            m2s_scatter_kernel<<<nb, 256, 0, st>>>(points, (long long)n, hist, xs, pm);
            err = (int)cudaGetLastError();
            pts = xs; perm = pm;
        }
    }
    if (!err && hier) {
        int* pbin = reinterpret_cast<int*>(ws + pbin_off);
        int* pidx = reinterpret_cast<int*>(ws + pidx_off);
        mesh2sdf_bound_kernel<<<(int)((n + M2S_WARP_THREADS - 1) / M2S_WARP_THREADS), M2S_WARP_THREADS, 0, st>>>(
            pts, (long long)n, recs, tflags, patches, (long long)num_tris, partial);
        const dim3 gd((unsigned)((n + M2S_WARP_THREADS - 1) / M2S_WARP_THREADS), (unsigned)dist_slices);
        mesh2sdf_dist_kernel<<<gd, M2S_WARP_THREADS, 0, st>>>(pts, (long long)n, recs, tsph, tflags, patches,
                                                              (long long)num_tris, partial);
        err = (int)cudaGetLastError();
        unsigned* ext = reinterpret_cast<unsigned*>(ws + ext_off);
        if (!err) err = (int)cudaMemsetAsync(pbin, 0, (size_t)proj_bins * 4, st);
        if (!err) err = (int)cudaMemsetAsync(ext, 0, 4, st);
        if (!err) {
            const int nb = (int)std::min<long long>((n + 255) / 256, sm8);
            m2s_extent_kernel<<<nb, 256, 0, st>>>(pts, (long long)n, tris, (long long)num_tris, ext);
            m2s_proj_bin_kernel<<<nb, 256, 0, st>>>(pts, (long long)n, G, ext, pbin, nullptr);
            m2s_scan(pbin, proj_bins, reinterpret_cast<int*>(ws + bsum_off), bsum_n, st);
            m2s_proj_bin_kernel<<<nb, 256, 0, st>>>(pts, (long long)n, G, ext, pbin, pidx);
            const long long tasks = (long long)num_tris * M2S_NDIR;
            const long long wpc = M2S_WARP_THREADS / 32;
            mesh2sdf_stab_kernel<<<(int)((tasks + wpc - 1) / wpc), M2S_WARP_THREADS, 0, st>>>(
                pts, recs, (long long)num_tris, G, ext, pbin, pidx, partial);
            err = (int)cudaGetLastError();
        }
    } else if (!err) {
        const dim3 g2((unsigned)grid, (unsigned)slices);
        mesh2sdf_kernel<<<g2, M2S_THREADS, 0, st>>>(pts, (long long)n, recs, (long long)num_tris, dist, perm, partial);
        err = (int)cudaGetLastError();
    }
    if (!err && partial) {
        m2s_finish_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(partial, (long long)n, dist, perm);
        err = (int)cudaGetLastError();
    }
    const int ferr = (int)cudaFreeAsync(ws, st);
    return err ? err : ferr;
}

extern "C" int nglod_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                               float lr, float beta1, float beta2, float eps, float bc1, float bc2, void* stream) {
    if (n < 0 || (n > 0 && (!param || !grad || !exp_avg || !exp_avg_sq))) return NGLOD_EINVAL;
    if (n == 0) return 0;
    long long grid = (n / 4 + 255) / 256 + 1;
    const long long cap = (long long)nglod_sm_count() * 16;
    if (grid > cap) grid = cap;
    adam_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, (long long)n, lr,
                                                             beta1, beta2, eps, bc1, sqrtf(bc2));
    return (int)cudaGetLastError();
}

extern "C" int nglod_abi_version(void) { return NGLOD_ABI_VERSION; }
extern "C" const char* nglod_build_info(void) {
#define NGLOD_STR2(x) #x
#define NGLOD_STR(x) NGLOD_STR2(x)
    return "nglod_b200 sm_100a (compute_100a) nvcc " NGLOD_STR(__CUDACC_VER_MAJOR__) "." NGLOD_STR(__CUDACC_VER_MINOR__) " built " __DATE__;
}
