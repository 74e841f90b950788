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

// rays are x-major: ray index = ix*H + iy (geoutils.py:149-153,190-194)
__global__ void __launch_bounds__(256)
generate_rays_kernel(const Camera cam, const float* __restrict__ wx, const float* __restrict__ wy, const int W,
                     const int H, float* __restrict__ ray_o, float* __restrict__ ray_d) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= (long long)W * H) return;
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
    for (int k = 0; k < 3; ++k) { ray_o[3 * i + k] = p[k]; ray_d[3 * i + k] = d[k] / den; }
}

// rgb = matcap(uv(view, normal)) / 255 on hits, 1 on misses; normals of misses are set to 1 (renderer.py:294-296)
__global__ void __launch_bounds__(256)
shade_matcap_kernel(const float* __restrict__ view, float* __restrict__ normal, const uint8_t* __restrict__ hit,
                    const float* __restrict__ tex, const int nu, const int nv, const int nc, const long long n,
                    float* __restrict__ rgb) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    if (!hit[i]) {
        normal[3 * i] = 1.f; normal[3 * i + 1] = 1.f; normal[3 * i + 2] = 1.f;
        rgb[3 * i] = 1.f; rgb[3 * i + 1] = 1.f; rgb[3 * i + 2] = 1.f;
        return;
    }
    const float nx = normal[3 * i], ny = normal[3 * i + 1], nz = normal[3 * i + 2];
    const float dx = view[3 * i], dy = view[3 * i + 1], dz = -view[3 * i + 2];
    const float dot = nx * dx + ny * dy + nz * dz;
    const float rx = dx - 2.0f * dot * nx, ry = dy - 2.0f * dot * ny, rz = dz - 2.0f * dot * nz - 1.0f;
    const float m = 2.0f * sqrtf(rx * rx + ry * ry + rz * rz);
    float u = 1.0f - (rx / m + 0.5f), v = 1.0f - (ry / m + 0.5f);
    u = fminf(1.f, fmaxf(0.f, u)); v = fminf(1.f, fmaxf(0.f, v));
    if (u != u) u = 0.f;
    if (v != v) v = 0.f;
    // bilinear on the [0,1]^2 lattice (RegularGridInterpolator 'linear')
    const float fu = u * (float)(nu - 1), fv = v * (float)(nv - 1);
    int u0 = min((int)floorf(fu), max(nu - 2, 0)), v0 = min((int)floorf(fv), max(nv - 2, 0));
    const int u1 = min(u0 + 1, nu - 1), v1 = min(v0 + 1, nv - 1);
    const float a = fu - (float)u0, b = fv - (float)v0;
    for (int c = 0; c < 3; ++c) {
        const float t00 = tex[((long long)u0 * nv + v0) * nc + c], t10 = tex[((long long)u1 * nv + v0) * nc + c];
        const float t01 = tex[((long long)u0 * nv + v1) * nc + c], t11 = tex[((long long)u1 * nv + v1) * nc + c];
        rgb[3 * i + c] = (t00 * (1 - a) * (1 - b) + t10 * a * (1 - b) + t01 * (1 - a) * b + t11 * a * b) * (1.0f / 255.0f);
    }
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
