/* A plain-C (C99) client of include/nglod_b200.h: what a maintainer's cgo / FFI stub sees.  Compiled and run by
 * tests/test_host_logic.py::test_header_is_plain_c_and_a_c_client_links -- it checks that the header needs no C++, that the
 * library links from C, the ABI version, and the entry points that do no device work (no GPU is needed to run it). */
#include <stdio.h>
#include <stddef.h>
#include <string.h>
#include <math.h>
#include "nglod_b200.h"

int main(void) {
    nglod_net_t net;
    nglod_net_grad_t grad;
    nglod_trace_opts_t opts;
    memset(&net, 0, sizeof net);
    memset(&grad, 0, sizeof grad);
    memset(&opts, 0, sizeof opts);
    if (nglod_abi_version() != NGLOD_ABI_VERSION) { printf("abi %d != header %d\n", nglod_abi_version(), NGLOD_ABI_VERSION); return 1; }
    const float from[3] = {-2.8f, 2.8f, -2.8f}, to[3] = {0.f, 0.f, 0.f};
    float basis[12];
    if (nglod_camera_basis(from, to, basis) != 0) return 2;
    /* view, right, up: unit length, mutually orthogonal, right has no y component (world up = +y) */
    for (int v = 1; v < 4; ++v) {
        const float* a = basis + 3 * v;
        if (fabsf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2] - 1.f) > 1e-5f) return 3;
    }
    if (fabsf(basis[3] * basis[6] + basis[4] * basis[7] + basis[5] * basis[8]) > 1e-6f || fabsf(basis[7]) > 1e-7f) return 4;
    /* argument checks that return before any CUDA call */
    if (nglod_camera_basis(NULL, to, basis) != NGLOD_EINVAL) return 5;
    if (nglod_sphere_trace(&net, 0, NULL, NULL, -1, &opts, NULL, NULL, NULL, NULL, NULL, NULL, NULL) == 0) return 6;
    if (nglod_sphere_trace_packed(&net, 0, NULL, NULL, 4, &opts, (float*)(void*)8, NULL, NULL, NULL, NULL) != NGLOD_EINVAL) return 7;
    printf("ok abi %d sizeof(nglod_net_t) %u sizeof(nglod_net_grad_t) %u sizeof(nglod_trace_opts_t) %u sizeof(nglod_sparse_net_t) %u "
           "build \"%s\"\n", nglod_abi_version(), (unsigned)sizeof net, (unsigned)sizeof grad, (unsigned)sizeof opts,
           (unsigned)sizeof(nglod_sparse_net_t), nglod_build_info());
    printf("offsets %u %u %u %u %u %u %u %u\n", (unsigned)offsetof(nglod_net_t, grid_res), (unsigned)offsetof(nglod_net_t, grids),
           (unsigned)offsetof(nglod_net_t, w0), (unsigned)offsetof(nglod_net_t, summed_fp16), (unsigned)offsetof(nglod_net_grad_t, summed),
           (unsigned)offsetof(nglod_net_grad_t, scatter_scratch_floats), (unsigned)offsetof(nglod_trace_opts_t, step_size),
           (unsigned)offsetof(nglod_trace_opts_t, normal_h));
    return 0;
}
