// nglod_b200 -- sparse-octree (SPC) ray traversal and first-voxel search.
//
// Behavioural spec: sol-renderer/include/spc/spc/spc_raytrace_cuda_kernel.cu:85-265 (the in-tree twin of Kaolin's
// unbatched_raytrace used by sdf-net/app/spc/SPC.py:104-107), sol-renderer/sdfRenderer.cu:108-120 (mark_first_hit) and
// sol-renderer/include/solr/solr/gfx/ray_aabb.cuh:42-192 (unbatched_ray_aabb).
//
// The reference walks the octree level-synchronously over ALL (ray, voxel) nuggets: per level a Decide kernel, a CUB
// scan, a Subdivide kernel and a device->host read of the count, on 2 x 2^27-entry zero-filled scratch buffers
// (2.1 GB per call).  Here every ray is walked depth-first by its own thread with a <=16-deep stack; because the
// reference expands nuggets in place, its output order per ray IS the pre-order leaf sequence, so the result is
// identical, nugget for nugget.  Two passes (count -> exclusive scan -> fill) replace the per-level scans: no scratch
// proportional to the nugget count, no host round trips inside the traversal (the caller reads ONE int, the total).
#include "common.cuh"

namespace {

constexpr int SPC_MAX_LEVELS = 16;
constexpr int SPC_THREADS = 128;
constexpr int SCAN_BLOCK = 1024;

struct SpcTree {
    const uint8_t* octree;      // child masks, breadth first
    const int32_t* prefix;      // exclusive sum of popcounts
    const short4* points;       // per-level Morton-ordered voxel coordinates
    int pyrsum[SPC_MAX_LEVELS + 2];
    int target;
};

// rows of the reference's front-to-back child order table (spc_raytrace_cuda_kernel.cu:39-47), 3 bits per entry
__constant__ unsigned c_order[8] = {
    0u | 1u << 3 | 2u << 6 | 4u << 9 | 3u << 12 | 5u << 15 | 6u << 18 | 7u << 21,
    1u | 0u << 3 | 3u << 6 | 5u << 9 | 2u << 12 | 4u << 15 | 7u << 18 | 6u << 21,
    2u | 0u << 3 | 3u << 6 | 6u << 9 | 1u << 12 | 4u << 15 | 7u << 18 | 5u << 21,
    3u | 1u << 3 | 2u << 6 | 7u << 9 | 0u << 12 | 5u << 15 | 6u << 18 | 4u << 21,
    4u | 0u << 3 | 5u << 6 | 6u << 9 | 1u << 12 | 2u << 15 | 7u << 18 | 3u << 21,
    5u | 1u << 3 | 4u << 6 | 7u << 9 | 0u << 12 | 3u << 15 | 6u << 18 | 2u << 21,
    6u | 2u << 3 | 4u << 6 | 7u << 9 | 0u << 12 | 3u << 15 | 5u << 18 | 1u << 21,
    7u | 3u << 3 | 5u << 6 | 6u << 9 | 1u << 12 | 2u << 15 | 4u << 18 | 0u << 21};

// d_FaceEval (:85-105) with the operation order nvcc emits for the reference source: r0 = fma(b,j,a*i) + c
__device__ __forceinline__ bool face_eval(float i, float j, float a, float b, float c) {
    const float r0 = __fadd_rn(__fmaf_rn(b, j, __fmul_rn(a, i)), c);
    const float r1 = __fadd_rn(r0, a), r2 = __fadd_rn(r0, b), r3 = __fadd_rn(r1, b);
    // the reference starts its running min / max at 1 / -1 and compares with `<` / `>` (a NaN never wins); the test below
    // only asks for the signs, for which min(1, r..) <= 0 <=> min(r..) <= 0 and fminf / fmaxf skip NaNs the same way:
    // 6 FMNMX instead of 16 compare-and-select (the traversal is instruction-bound, profiles/README.md part 4)
    const float mn = fminf(fminf(r0, r1), fminf(r2, r3)), mx = fmaxf(fmaxf(r0, r1), fmaxf(r2, r3));
    return mn <= 0.0f && mx >= 0.0f;
}

struct RayPre { float ox, oy, oz, dx, dy, dz, cx, cy, cz; };

// d_Decide (:108-138): infinite-line vs voxel overlap from three projected tests
__device__ __forceinline__ bool decide(const RayPre& r, short4 p, int level) {
    const float s1 = __int_as_float((127 - level) << 23), s2 = s1 * s1;      // 2^-level, exactly what 1.0f / (1 << level) gives
    const float px = (float)(unsigned short)p.x, py = (float)(unsigned short)p.y, pz = (float)(unsigned short)p.z;
    return face_eval(py, pz, -s2 * r.dz, s2 * r.dy, s1 * r.cx) && face_eval(px, pz, s2 * r.dz, -s2 * r.dx, s1 * r.cy) &&
           face_eval(px, py, -s2 * r.dy, s2 * r.dx, s1 * r.cz);
}

__device__ __forceinline__ RayPre ray_pre(const float* __restrict__ ray_o, const float* __restrict__ ray_d, int ray) {
    RayPre r;
    const float o0 = __ldg(ray_o + 3 * ray), o1 = __ldg(ray_o + 3 * ray + 1), o2 = __ldg(ray_o + 3 * ray + 2);
    r.ox = __fmaf_rn(0.5f, o0, 0.5f); r.oy = __fmaf_rn(0.5f, o1, 0.5f); r.oz = __fmaf_rn(0.5f, o2, 0.5f);
    r.dx = 0.5f * __ldg(ray_d + 3 * ray); r.dy = 0.5f * __ldg(ray_d + 3 * ray + 1); r.dz = 0.5f * __ldg(ray_d + 3 * ray + 2);
    r.cx = __fmaf_rn(r.oy, r.dz, -__fmul_rn(r.dy, r.oz));
    r.cy = __fmaf_rn(r.oz, r.dx, -__fmul_rn(r.dz, r.ox));
    r.cz = __fmaf_rn(r.ox, r.dy, -__fmul_rn(r.dx, r.oy));
    return r;
}

// Depth-first walk of one ray; emit(pidx) is called for every leaf of the target level the ray's line overlaps, in the
// reference's order (front to back).
template <typename Emit>
__device__ __forceinline__ void spc_walk(const SpcTree& tree, const RayPre& r, Emit emit) {
    // per-level frame: child-prefix of the node, {mask, remaining order nibbles}, children still to try
    int f_s[SPC_MAX_LEVELS];
    unsigned f_state[SPC_MAX_LEVELS];       // bits 0-7 mask, 8-31 remaining order nibbles (3 bits each), consumed from the low end
    unsigned char f_left[SPC_MAX_LEVELS];   // children still to try
    auto push = [&](int level, int g, short4 p) {     // node (level, g) passed Decide and is above the target level
        const unsigned mask = __ldg(tree.octree + g);
        f_s[level] = __ldg(tree.prefix + g);
        const float scale = __int_as_float((127 - level) << 23);
        // octant of the ray origin relative to the voxel centre (:160-167); exact in double like the reference
        const double x = (double)r.ox - (double)scale * ((double)(unsigned short)p.x + 0.5);
        const double y = (double)r.oy - (double)scale * ((double)(unsigned short)p.y + 0.5);
        const double z = (double)r.oz - (double)scale * ((double)(unsigned short)p.z + 0.5);
        int code = 0;
        if ((float)x > 0.f) code = 4;
        if ((float)y > 0.f) code += 2;
        if ((float)z > 0.f) code += 1;
        f_state[level] = mask | (c_order[code] << 8);
        f_left[level] = 8;
    };
    const short4 proot = __ldg(tree.points);
    int sp = -1;
    if (decide(r, proot, 0)) {
        if (tree.target == 0) emit(0);
        else { push(0, 0, proot); sp = 0; }
    }
    while (sp >= 0) {
        if (f_left[sp] == 0) { --sp; continue; }
        const unsigned st = f_state[sp];
        const unsigned j = (st >> 8) & 7u;
        f_state[sp] = (st & 0xFFu) | ((st >> 11) << 8);
        --f_left[sp];
        if (!((st >> j) & 1u)) continue;
        const int cnt = __popc(st & 0xFFu & ((2u << j) - 1u));
        const int g = f_s[sp] + cnt;                 // global index of the child (root is 0)
        const int level = sp + 1;
        const short4 p = __ldg(tree.points + g);
        if (!decide(r, p, level)) continue;
        if (level == tree.target) emit(g - tree.pyrsum[level]);
        else { push(level, g, p); sp = level; }
    }
}

template <bool WRITE>
__global__ void __launch_bounds__(SPC_THREADS)
spc_raytrace_kernel(const SpcTree tree, const float* __restrict__ ray_o, const float* __restrict__ ray_d, const int n,
                    int* __restrict__ counts, const int* __restrict__ offsets, int2* __restrict__ nuggets) {
    const int ray = blockIdx.x * SPC_THREADS + threadIdx.x;
    if (ray >= n) return;
    const RayPre r = ray_pre(ray_o, ray_d, ray);
    int count = 0;
    int wpos = WRITE ? offsets[ray] : 0;
    spc_walk(tree, r, [&](int pidx) {
        if (WRITE) nuggets[wpos++] = make_int2(ray, pidx);
        ++count;
    });
    if (!WRITE) counts[ray] = count;
}

// One pass for a consumer that only needs every ray's OWN run (the in-voxel tracer): walk, keep the first RUN_BUF leaves
// in registers, reserve the run with one atomicAdd per warp (runs of different rays land in arbitrary order; inside a
// run the order is the reference's), write from the buffer -- or walk again when the run is longer.  No scan, no second
// launch, no host read of the total; cursor[1] is set when `capacity` is too small (the caller then takes the exact
// two-pass path).  The average run of a 1080p frame over a level-7 shell is 2.2 nuggets, the longest a few dozen.
constexpr int RUN_BUF = 8;
__global__ void __launch_bounds__(SPC_THREADS)
spc_raytrace_runs_kernel(const SpcTree tree, const float* __restrict__ ray_o, const float* __restrict__ ray_d, const int n,
                         const int capacity, int2* __restrict__ nuggets, int* __restrict__ run_begin,
                         int* __restrict__ run_end, int* __restrict__ cursor) {
    const int ray = blockIdx.x * SPC_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool live = ray < n;
    RayPre r = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int count = 0;
    int buf[RUN_BUF];
    if (live) {
        r = ray_pre(ray_o, ray_d, ray);
        spc_walk(tree, r, [&](int pidx) {
#pragma unroll
            for (int k = 0; k < RUN_BUF; ++k)
                if (k == count) buf[k] = pidx;
            ++count;
        });
    }
    // warp-aggregated reservation
    int incl = count;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int base = 0;
    if (lane == 31 && total > 0) base = atomicAdd(cursor, total);
    base = __shfl_sync(0xffffffffu, base, 31);
    if (!live) return;
    int begin = base + incl - count;
    if (count > 0 && (long long)begin + count > (long long)capacity) {
        atomicExch(cursor + 1, 1);
        count = 0;
    }
    if (count == 0) begin = 0;
    run_begin[ray] = begin;
    run_end[ray] = begin + count;
    if (count <= RUN_BUF) {
#pragma unroll
        for (int k = 0; k < RUN_BUF; ++k)
            if (k < count) nuggets[begin + k] = make_int2(ray, buf[k]);
    } else {
        int wpos = begin;
        spc_walk(tree, r, [&](int pidx) { nuggets[wpos++] = make_int2(ray, pidx); });
    }
}

// ---- exclusive scan of int32 counts (n up to 2^31): block scan -> scan of block sums -> add
__global__ void __launch_bounds__(SCAN_BLOCK)
scan_blocks_kernel(const int* __restrict__ in, int* __restrict__ out, int* __restrict__ block_sums, const int n) {
    __shared__ int warp_sums[32];
    const int i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int v = i < n ? in[i] : 0;
    int s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
    if (lane == 31) warp_sums[warp] = s;
    __syncthreads();
    if (warp == 0) {
        int w = warp_sums[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
        warp_sums[lane] = w;
    }
    __syncthreads();
    const int incl = s + (warp ? warp_sums[warp - 1] : 0);
    if (i < n) out[i] = incl - v;
    if (threadIdx.x == SCAN_BLOCK - 1) block_sums[blockIdx.x] = incl;
}

__global__ void __launch_bounds__(SCAN_BLOCK)
scan_sums_kernel(int* __restrict__ block_sums, const int nblocks, int* __restrict__ total_out) {
    __shared__ int warp_sums[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < nblocks; base += SCAN_BLOCK) {
        const int i = base + threadIdx.x;
        const int v = i < nblocks ? block_sums[i] : 0;
        int s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
        if (lane == 31) warp_sums[warp] = s;
        __syncthreads();
        if (warp == 0) {
            int w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const int incl = s + (warp ? warp_sums[warp - 1] : 0) + carry;
        if (i < nblocks) block_sums[i] = incl - v;          // exclusive
        __syncthreads();
        if (threadIdx.x == SCAN_BLOCK - 1) carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(SCAN_BLOCK)
scan_add_kernel(int* __restrict__ out, const int* __restrict__ block_sums, const int n) {
    const int i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    if (i < n) out[i] += block_sums[blockIdx.x];
}

__global__ void __launch_bounds__(256)
mark_first_hit_kernel(const int2* __restrict__ nuggets, const long long m, int* __restrict__ info) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i < m) info[i] = (i == 0 || nuggets[i - 1].x != nuggets[i].x) ? 1 : 0;
}

// device primitive of ray_aabb.cuh:42-89
__device__ __forceinline__ float ray_aabb_voxel(float qx, float qy, float qz, float dx, float dy, float dz, float ix,
                                                float iy, float iz, float sx, float sy, float sz, float vx, float vy,
                                                float vz, float r) {
    const float ox = qx - vx, oy = qy - vy, oz = qz - vz;
    const float cmax = fmaxf(fmaxf(fabsf(ox), fabsf(oy)), fabsf(oz));
    float winding = cmax < r ? -1.0f : 1.0f;
    winding *= r;
    if (winding < 0.f) return winding;
    const float d0 = __fmul_rn(__fmaf_rn(winding, sx, -ox), ix);
    const float d1 = __fmul_rn(__fmaf_rn(winding, sy, -oy), iy);
    const float d2 = __fmul_rn(__fmaf_rn(winding, sz, -oz), iz);
    const float ltxy = __fmaf_rn(dy, d0, oy), ltxz = __fmaf_rn(dz, d0, oz);
    const float ltyx = __fmaf_rn(dx, d1, ox), ltyz = __fmaf_rn(dz, d1, oz);
    const float ltzx = __fmaf_rn(dx, d2, ox), ltzy = __fmaf_rn(dy, d2, oy);
    if ((d0 >= 0.0f) && (fabsf(ltxy) < r) && (fabsf(ltxz) < r)) return d0;
    if ((d1 >= 0.0f) && (fabsf(ltyx) < r) && (fabsf(ltyz) < r)) return d1;
    if ((d2 >= 0.0f) && (fabsf(ltzx) < r) && (fabsf(ltzy) < r)) return d2;
    return 0.0f;
}

// One thread per ray: walk the ray's nugget run [offsets[ray], offsets[ray+1]) front to back (ray_aabb.cuh:104-192).
__global__ void __launch_bounds__(256)
spc_ray_aabb_kernel(const int2* __restrict__ nuggets, const int* __restrict__ offsets, const int n_rays,
                    const short4* __restrict__ level_points, const float r, const float* __restrict__ ray_o,
                    const float* __restrict__ ray_d, const float* __restrict__ query,
                    const uint8_t* __restrict__ active, float* __restrict__ x, float* __restrict__ t,
                    uint8_t* __restrict__ cond, int* __restrict__ pidx) {
    const int ray = blockIdx.x * 256 + threadIdx.x;
    if (ray >= n_rays) return;
    const int beg = offsets[ray], end = offsets[ray + 1];
    if (beg == end) return;                          // no run: the reference never touches such rays
    if (active && !active[ray]) return;              // `if (!cond[ridx] && !init) continue;`
    const float dx = ray_d[3 * ray], dy = ray_d[3 * ray + 1], dz = ray_d[3 * ray + 2];
    const float ix = 1.0f / dx, iy = 1.0f / dy, iz = 1.0f / dz;
    const float sx = signbit(dx) ? 1.0f : -1.0f, sy = signbit(dy) ? 1.0f : -1.0f, sz = signbit(dz) ? 1.0f : -1.0f;
    const float qx = query[3 * ray], qy = query[3 * ray + 1], qz = query[3 * ray + 2];
    bool hit = false;
    for (int i = beg; i < end && !hit; ++i) {
        const int pi = nuggets[i].y;
        const short4 p = __ldg(level_points + pi);
        const float vx = __fmaf_rn(r, __fmaf_rn(2.0f, (float)p.x, 1.0f), -1.0f);
        const float vy = __fmaf_rn(r, __fmaf_rn(2.0f, (float)p.y, 1.0f), -1.0f);
        const float vz = __fmaf_rn(r, __fmaf_rn(2.0f, (float)p.z, 1.0f), -1.0f);
        const float d = ray_aabb_voxel(qx, qy, qz, dx, dy, dz, ix, iy, iz, sx, sy, sz, vx, vy, vz, r);
        if (d != 0.0f) {
            hit = true;
            pidx[ray] = pi;
            cond[ray] = 1;
            if (d > 0.0f) {
                const float tt = t[ray] + d;
                t[ray] = tt;
                x[3 * ray] = __fmaf_rn(dx, tt, ray_o[3 * ray]);
                x[3 * ray + 1] = __fmaf_rn(dy, tt, ray_o[3 * ray + 1]);
                x[3 * ray + 2] = __fmaf_rn(dz, tt, ray_o[3 * ray + 2]);
            }
        }
    }
    if (!hit) {
        cond[ray] = 0;
        t[ray] = 100.0f;
        x[3 * ray] = __fmaf_rn(dx, 100.0f, ray_o[3 * ray]);
        x[3 * ray + 1] = __fmaf_rn(dy, 100.0f, ray_o[3 * ray + 1]);
        x[3 * ray + 2] = __fmaf_rn(dz, 100.0f, ray_o[3 * ray + 2]);
    }
}

int make_tree(SpcTree& t, const uint8_t* octree, const int32_t* prefix, const int16_t* points, const int32_t* pyramid_sum,
              int level, int target) {
    if (!octree || !prefix || !points || !pyramid_sum) return NGLOD_EINVAL;
    if (level < 0 || level >= SPC_MAX_LEVELS || target < 0 || target > level) return NGLOD_EINVAL;
    if (reinterpret_cast<uintptr_t>(points) & 7u) return NGLOD_EINVAL;
    t.octree = octree; t.prefix = prefix; t.points = reinterpret_cast<const short4*>(points); t.target = target;
    for (int i = 0; i < SPC_MAX_LEVELS + 2; ++i) t.pyrsum[i] = i <= level + 1 ? pyramid_sum[i] : 0;
    return 0;
}

}  // namespace

extern "C" int nglod_spc_raytrace_count(const uint8_t* octree, const int32_t* prefix, const int16_t* points,
                                        const int32_t* pyramid_sum, int32_t level, int32_t target_level,
                                        const float* ray_o, const float* ray_d, int64_t n, int32_t* offsets,
                                        int32_t* scan_ws, void* stream) {
    SpcTree tree;
    if (int e = make_tree(tree, octree, prefix, points, pyramid_sum, level, target_level)) return e;
    if (n < 0 || n > 2000000000ll || !offsets || !scan_ws || (n > 0 && (!ray_o || !ray_d))) return NGLOD_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) return (int)cudaMemsetAsync(offsets, 0, sizeof(int32_t), st);
    const int nb = (int)((n + SPC_THREADS - 1) / SPC_THREADS);
    spc_raytrace_kernel<false><<<nb, SPC_THREADS, 0, st>>>(tree, ray_o, ray_d, (int)n, offsets, nullptr, nullptr);
    const int sb = (int)((n + SCAN_BLOCK - 1) / SCAN_BLOCK);
    scan_blocks_kernel<<<sb, SCAN_BLOCK, 0, st>>>(offsets, offsets, scan_ws, (int)n);
    scan_sums_kernel<<<1, SCAN_BLOCK, 0, st>>>(scan_ws, sb, offsets + n);
    scan_add_kernel<<<sb, SCAN_BLOCK, 0, st>>>(offsets, scan_ws, (int)n);
    return (int)cudaGetLastError();
}

extern "C" int nglod_spc_raytrace_fill(const uint8_t* octree, const int32_t* prefix, const int16_t* points,
                                       const int32_t* pyramid_sum, int32_t level, int32_t target_level,
                                       const float* ray_o, const float* ray_d, int64_t n, const int32_t* offsets,
                                       int32_t* nuggets, void* stream) {
    SpcTree tree;
    if (int e = make_tree(tree, octree, prefix, points, pyramid_sum, level, target_level)) return e;
    if (n < 0 || n > 2000000000ll || !offsets || (n > 0 && (!ray_o || !ray_d))) return NGLOD_EINVAL;
    if (n == 0) return 0;
    if (!nuggets || (reinterpret_cast<uintptr_t>(nuggets) & 7u)) return NGLOD_EINVAL;
    const int nb = (int)((n + SPC_THREADS - 1) / SPC_THREADS);
    spc_raytrace_kernel<true><<<nb, SPC_THREADS, 0, (cudaStream_t)stream>>>(tree, ray_o, ray_d, (int)n, nullptr, offsets,
                                                                            reinterpret_cast<int2*>(nuggets));
    return (int)cudaGetLastError();
}

extern "C" int nglod_spc_raytrace_runs(const uint8_t* octree, const int32_t* prefix, const int16_t* points,
                                       const int32_t* pyramid_sum, int32_t level, int32_t target_level,
                                       const float* ray_o, const float* ray_d, int64_t n, int64_t capacity,
                                       int32_t* nuggets, int32_t* run_begin, int32_t* run_end, int32_t* cursor, void* stream) {
    SpcTree tree;
    if (int e = make_tree(tree, octree, prefix, points, pyramid_sum, level, target_level)) return e;
    if (n < 0 || n > 2000000000ll || capacity < 0 || capacity > 2000000000ll || !cursor) return NGLOD_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    NGLOD_CUDA_TRY(cudaMemsetAsync(cursor, 0, 2 * sizeof(int32_t), st));
    if (n == 0) return 0;
    if (!ray_o || !ray_d || !run_begin || !run_end || (capacity > 0 && !nuggets)) return NGLOD_EINVAL;
    if (reinterpret_cast<uintptr_t>(nuggets) & 7u) return NGLOD_EINVAL;
    const int nb = (int)((n + SPC_THREADS - 1) / SPC_THREADS);
    spc_raytrace_runs_kernel<<<nb, SPC_THREADS, 0, st>>>(tree, ray_o, ray_d, (int)n, (int)capacity,
                                                         reinterpret_cast<int2*>(nuggets), run_begin, run_end, cursor);
    return (int)cudaGetLastError();
}

extern "C" int nglod_spc_mark_first_hit(const int32_t* nuggets, int64_t m, int32_t* info, void* stream) {
    if (m < 0 || (m > 0 && (!nuggets || !info))) return NGLOD_EINVAL;
    if (m == 0) return 0;
    mark_first_hit_kernel<<<(unsigned)((m + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const int2*>(nuggets), (long long)m, info);
    return (int)cudaGetLastError();
}

extern "C" int nglod_spc_ray_aabb(const int32_t* nuggets, const int32_t* offsets, int64_t n_rays,
                                  const int16_t* level_points, int32_t level, const float* ray_o, const float* ray_d,
                                  const float* query, const uint8_t* active, float* x, float* t, uint8_t* cond,
                                  int32_t* pidx, void* stream) {
    if (n_rays < 0 || n_rays > 2000000000ll) return NGLOD_EINVAL;
    if (n_rays == 0) return 0;
    if (!offsets || !level_points || !ray_o || !ray_d || !query || !x || !t || !cond || !pidx) return NGLOD_EINVAL;
    if (level < 0 || level >= SPC_MAX_LEVELS) return NGLOD_EINVAL;
    const float r = 1.0f / (float)(1 << level);
    spc_ray_aabb_kernel<<<(unsigned)((n_rays + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const int2*>(nuggets), offsets, (int)n_rays, reinterpret_cast<const short4*>(level_points), r,
        ray_o, ray_d, query, active, x, t, cond, pidx);
    return (int)cudaGetLastError();
}
