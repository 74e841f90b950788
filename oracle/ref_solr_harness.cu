// TEST INFRASTRUCTURE -- C-ABI launchers around the reference's OWN sol-renderer kernels, which are #included
// unmodified from where they lie (-I /root/reference/sol-renderer/include/solr; nothing is copied into this repo):
//   solr/gfx/ray_aabb.cuh:104-192          ray_aabb_kernel (nugget walk: first voxel containing / entered from `query`)
//   solr/sdf/sparse_grid_sample.cuh:31-109 sparse_grid_sample_kernel (trinkets + parent chain -> [x, features])
//   solr/sdf/step.cuh:31-86                step_kernel (t += d, hit / cond update, x = o + t d)
//   solr/sdf/index_trinket.cuh:30-99       index_trinket_kernel (trinkets + parents from points and corner coords)
//   solr/common/normalize.cuh:28-47        normalize_kernel
// Launch shapes follow the reference's call sites in sol-renderer/SDF.cu (:190-205, :257-284, :355-372, :404-462).
// Built by oracle/build_ref.py into oracle/_ref/ref_solr/ref_solr.so; loaded with ctypes by tests/test_spc_ref.py.
#include <cuda_runtime.h>
#include <stdint.h>
#include "solr/gfx/ray_aabb.cuh"
#include "solr/sdf/sparse_grid_sample.cuh"
#include "solr/sdf/step.cuh"
#include "solr/sdf/index_trinket.cuh"
#include "solr/common/normalize.cuh"

static int done() {
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaGetLastError();
    return (int)e;
}

extern "C" int ref_solr_sizeof_trinket() { return (int)sizeof(solr::Trinket); }

extern "C" int ref_solr_ray_aabb(const float* ray_o, const float* ray_d, const float* ray_inv, const float* query,
                                 const void* nuggets, const short* points, const int* info, const int* info_idxes,
                                 float r, int init, float* x, float* t, bool* cond, int* pidx, int num_nuggets,
                                 int n_idx, int threads) {
    if (n_idx <= 0) return 0;
    const int blocks = (n_idx + threads - 1) / threads;
    solr::ray_aabb_kernel<<<blocks, threads>>>(ray_o, ray_d, ray_inv, query, reinterpret_cast<const solr::Nugget*>(nuggets),
                                               points, info, info_idxes, r, init != 0, x, t, cond, pidx, num_nuggets, n_idx);
    return done();
}

extern "C" int ref_solr_sparse_grid_sample(const float* x, const int* pidx, const int* idxes, const void* trinkets,
                                           const float* feats_in, const unsigned* pyramid, const unsigned* resolutions,
                                           float* feats_out, int n, int m, int dim, int nl, int lod, const int* cc) {
    if (n <= 0) return 0;
    const int threads = 128, blocks = (n + threads - 1) / threads;
    solr::sparse_grid_sample_kernel<<<blocks, threads>>>(x, pidx, idxes, reinterpret_cast<const solr::Trinket*>(trinkets),
                                                         feats_in, pyramid, resolutions, feats_out, n, m, dim, nl, lod, cc);
    return done();
}

extern "C" int ref_solr_step(const float* ray_o, const float* ray_d, const int* idxes, const float* d_in, float* x, float* t,
                             float* d, float* dprev, bool* cond, bool* hit, int n) {
    if (n <= 0) return 0;
    const int threads = 128, blocks = (n + threads - 1) / threads;
    solr::step_kernel<<<blocks, threads>>>(ray_o, ray_d, idxes, d_in, x, t, d, dprev, cond, hit, n);
    return done();
}

extern "C" int ref_solr_normalize(const int* idxes, float* x, int n) {
    if (n <= 0) return 0;
    const int threads = 128, blocks = (n + threads - 1) / threads;
    solr::normalize_kernel<<<blocks, threads>>>(idxes, x, n);
    return done();
}

extern "C" int ref_solr_index_trinkets(const void* points, const int* coords, const float* feats, void* trinkets,
                                       int num_coords, int offset_cf, int n, int offset, int parent_n, int parent_offset,
                                       int level) {
    if (n <= 0) return 0;
    const int threads = 1024, blocks = (n + threads - 1) / threads;
    solr::index_trinket_kernel<<<blocks, threads>>>(reinterpret_cast<const ushort4*>(points), coords, feats,
                                                    reinterpret_cast<solr::Trinket*>(trinkets), num_coords, offset_cf, n, offset,
                                                    parent_n, parent_offset, level);
    return done();
}
