// nglod_b200 -- small device-side pieces of the renderer entry point so a frame never bounces through the host.
//
//   generate_rays  <- look_at + normalized_grid, sdf-net/lib/geoutils.py:140-154,180-206 (the per-column / per-row
//                     jitter is drawn by the caller with torch.rand exactly like the reference, so seeded runs agree)
//   shade_matcap   <- spherical_envmap + matcap lookup + "misses are white", sdf-net/lib/geoutils.py:253-275 and
//                     lib/renderer.py:279-296 (the reference copies UVs to the host, interpolates with scipy's
//                     RegularGridInterpolator and copies colours back; here it is one kernel)
#include "common.cuh"

namespace {

struct Camera {
    float origin[3], view[3], right[3], up[3];
    float tan_half_fov;
    int ortho;
};

// [N,3] fp32 rows are 12 bytes apart: a warp that reads or writes its 32 rows with scalar accesses touches every 32-byte
// sector three times with a third of it each.  The kernels below move a block's 256 rows (3072 contiguous bytes per array)
// as 192 16-byte vectors through shared memory instead (stride-3 shared accesses are conflict-free); rows of a partial
// last block, or arrays that are not 16-byte aligned (a slice starting at an odd row), take the scalar path.
constexpr int RB = 256;                       // rows per block
__device__ __forceinline__ bool rows_vectorisable(const void* p, long long row0, long long n) {
    return (reinterpret_cast<uintptr_t>(p) & 15u) == 0 && row0 + RB <= n;
}
__device__ __forceinline__ void rows_load(float* sm, const float* __restrict__ g, long long row0, long long n, bool vec) {
    if (vec) {
        if (threadIdx.x < RB * 3 / 4)
            reinterpret_cast<float4*>(sm)[threadIdx.x] = __ldg(reinterpret_cast<const float4*>(g + 3 * row0) + threadIdx.x);
    } else if (row0 + threadIdx.x < n) {
        const float* r = g + 3 * (row0 + threadIdx.x);
        sm[3 * threadIdx.x] = r[0]; sm[3 * threadIdx.x + 1] = r[1]; sm[3 * threadIdx.x + 2] = r[2];
    }
}
__device__ __forceinline__ void rows_store(const float* sm, float* __restrict__ g, long long row0, long long n, bool vec) {
    if (vec) {
        if (threadIdx.x < RB * 3 / 4)
            reinterpret_cast<float4*>(g + 3 * row0)[threadIdx.x] = reinterpret_cast<const float4*>(sm)[threadIdx.x];
    } else if (row0 + threadIdx.x < n) {
        float* r = g + 3 * (row0 + threadIdx.x);
        r[0] = sm[3 * threadIdx.x]; r[1] = sm[3 * threadIdx.x + 1]; r[2] = sm[3 * threadIdx.x + 2];
    }
}

// rays are x-major: ray index = ix*H + iy (geoutils.py:149-153,190-194)
__global__ void __launch_bounds__(RB)
generate_rays_kernel(const Camera cam, const float* __restrict__ wx, const float* __restrict__ wy, const int W,
                     const int H, float* __restrict__ ray_o, float* __restrict__ ray_d) {
    __shared__ __align__(16) float so[RB * 3], sd[RB * 3];
    const long long n = (long long)W * H;
    const long long row0 = (long long)blockIdx.x * RB;
    const long long i = row0 + threadIdx.x;
    if (i < n) {
        const int ix = (int)(i / H), iy = (int)(i - (long long)ix * H);
        const float cx = wx[ix], cy = wy[iy];
        float p[3], d[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            // right*cx*tan + up*cy*tan + origin + view, left to right like the reference expression
            const float a = __fmul_rn(__fmul_rn(cam.right[k], cx), cam.tan_half_fov);
            const float b = __fmul_rn(__fmul_rn(cam.up[k], cy), cam.tan_half_fov);
            p[k] = __fadd_rn(__fadd_rn(__fadd_rn(a, b), cam.origin[k]), cam.view[k]);
        }
        if (cam.ortho) {
#pragma unroll
            for (int k = 0; k < 3; ++k) d[k] = cam.view[k];
        } else {
#pragma unroll
            for (int k = 0; k < 3; ++k) { d[k] = __fsub_rn(p[k], cam.origin[k]); p[k] = cam.origin[k]; }
        }
        // F.normalize(dim=-1): v / max(||v||, 1e-12)
        const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
        const float den = fmaxf(nrm, 1e-12f);
#pragma unroll
        for (int k = 0; k < 3; ++k) { so[3 * threadIdx.x + k] = p[k]; sd[3 * threadIdx.x + k] = d[k] / den; }
    }
    __syncthreads();
    rows_store(so, ray_o, row0, n, rows_vectorisable(ray_o, row0, n));
    rows_store(sd, ray_d, row0, n, rows_vectorisable(ray_d, row0, n));
}

// rgb = matcap(uv(view, normal)) / 255 on hits, 1 on misses; normals of misses are set to 1 (renderer.py:294-296)
__global__ void __launch_bounds__(RB)
shade_matcap_kernel(const float* __restrict__ view, float* __restrict__ normal, const uint8_t* __restrict__ hit,
                    const float* __restrict__ tex, const int nu, const int nv, const int nc, const long long n,
                    float* __restrict__ rgb) {
    __shared__ __align__(16) float sv[RB * 3], sn[RB * 3];      // view -> rgb in place, normal in place
    const long long row0 = (long long)blockIdx.x * RB;
    const long long i = row0 + threadIdx.x;
    const bool vec_v = rows_vectorisable(view, row0, n), vec_n = rows_vectorisable(normal, row0, n);
    rows_load(sv, view, row0, n, vec_v);
    rows_load(sn, normal, row0, n, vec_n);
    __syncthreads();
    if (i < n) {
        float* v = sv + 3 * threadIdx.x;
        float* nm = sn + 3 * threadIdx.x;
        if (!hit[i]) {
            nm[0] = 1.f; nm[1] = 1.f; nm[2] = 1.f;
            v[0] = 1.f; v[1] = 1.f; v[2] = 1.f;
        } else {
            const float nx = nm[0], ny = nm[1], nz = nm[2];
            const float dx = v[0], dy = v[1], dz = -v[2];
            const float dot = nx * dx + ny * dy + nz * dz;
            const float rx = dx - 2.0f * dot * nx, ry = dy - 2.0f * dot * ny, rz = dz - 2.0f * dot * nz - 1.0f;
            const float m = 2.0f * sqrtf(rx * rx + ry * ry + rz * rz);
            float u = 1.0f - (rx / m + 0.5f), w = 1.0f - (ry / m + 0.5f);
            u = fminf(1.f, fmaxf(0.f, u)); w = fminf(1.f, fmaxf(0.f, w));
            if (u != u) u = 0.f;
            if (w != w) w = 0.f;
            // bilinear on the [0,1]^2 lattice (RegularGridInterpolator 'linear')
            const float fu = u * (float)(nu - 1), fv = w * (float)(nv - 1);
            const int u0 = min((int)floorf(fu), max(nu - 2, 0)), v0 = min((int)floorf(fv), max(nv - 2, 0));
            const int u1 = min(u0 + 1, nu - 1), v1 = min(v0 + 1, nv - 1);
            const float a = fu - (float)u0, b = fv - (float)v0;
            for (int c = 0; c < 3; ++c) {
                const float t00 = tex[((long long)u0 * nv + v0) * nc + c], t10 = tex[((long long)u1 * nv + v0) * nc + c];
                const float t01 = tex[((long long)u0 * nv + v1) * nc + c], t11 = tex[((long long)u1 * nv + v1) * nc + c];
                v[c] = (t00 * (1 - a) * (1 - b) + t10 * a * (1 - b) + t01 * (1 - a) * b + t11 * a * b) * (1.0f / 255.0f);
            }
        }
    }
    __syncthreads();
    rows_store(sv, rgb, row0, n, rows_vectorisable(rgb, row0, n));
    rows_store(sn, normal, row0, n, vec_n);
}

}  // namespace

// Host-only: the camera frame of look_at (geoutils.py:180-188) in the float32 arithmetic torch's CPU kernels use for
// 3-vectors -- norm = sqrt(fma(z, z, fma(y, y, x * x))), cross component = fma(a1, b2, -(a2 * b1)), F.normalize's
// v / max(norm, 1e-12) -- so that the basis is the one `F.normalize(torch.linalg.cross(...))` returns on the host bit
// for bit (tests/test_host_logic.py pins that on random poses) at 1/30 of the cost of eight torch calls.
static void basis_normalize(float* v) {
    const float n = sqrtf(fmaf(v[2], v[2], fmaf(v[1], v[1], v[0] * v[0])));
    const float d = n > 1e-12f ? n : 1e-12f;
    v[0] = v[0] / d; v[1] = v[1] / d; v[2] = v[2] / d;
}
static void basis_cross(const float* a, const float* b, float* c) {
    const float p0 = a[2] * b[1], p1 = a[0] * b[2], p2 = a[1] * b[0];
    c[0] = fmaf(a[1], b[2], -p0);
    c[1] = fmaf(a[2], b[0], -p1);
    c[2] = fmaf(a[0], b[1], -p2);
}
extern "C" int nglod_camera_basis(const float* from, const float* to, float* basis) {
    if (!from || !to || !basis) return NGLOD_EINVAL;
    float* origin = basis; float* view = basis + 3; float* right = basis + 6; float* up = basis + 9;
    const float world_up[3] = {0.f, 1.f, 0.f};
    for (int k = 0; k < 3; ++k) { origin[k] = from[k]; view[k] = to[k] - from[k]; }
    basis_normalize(view);
    basis_cross(view, world_up, right);
    basis_normalize(right);
    basis_cross(right, view, up);
    basis_normalize(up);
    return 0;
}

extern "C" int nglod_generate_rays(const float* origin, const float* view, const float* right, const float* up,
                                   float tan_half_fov, int32_t ortho, const float* window_x, const float* window_y,
                                   int32_t width, int32_t height, float* ray_o, float* ray_d, void* stream) {
    if (!origin || !view || !right || !up || width < 0 || height < 0) return NGLOD_EINVAL;
    const long long n = (long long)width * height;
    if (n == 0) return 0;
    if (!window_x || !window_y || !ray_o || !ray_d) return NGLOD_EINVAL;
    Camera cam;
    for (int k = 0; k < 3; ++k) { cam.origin[k] = origin[k]; cam.view[k] = view[k]; cam.right[k] = right[k]; cam.up[k] = up[k]; }
    cam.tan_half_fov = tan_half_fov;
    cam.ortho = ortho;
    generate_rays_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(cam, window_x, window_y, width,
                                                                                       height, ray_o, ray_d);
    return (int)cudaGetLastError();
}

extern "C" int nglod_shade_matcap(const float* view, float* normal, const uint8_t* hit, const float* matcap,
                                  int32_t nu, int32_t nv, int32_t nc, int64_t n, float* rgb, void* stream) {
    if (n < 0 || nu < 1 || nv < 1 || nc < 3) return NGLOD_EINVAL;
    if (n == 0) return 0;
    if (!view || !normal || !hit || !matcap || !rgb) return NGLOD_EINVAL;
    shade_matcap_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(view, normal, hit, matcap, nu, nv,
                                                                                      nc, (long long)n, rgb);
    return (int)cudaGetLastError();
}
