/*
 * nglod_b200 -- C ABI of the B200-native NGLOD hot path.
 *
 * One shared library (libnglod_b200.so), every entry point `extern "C"`, plain
 * device pointers and sizes, an explicit cudaStream_t (passed as void*), no
 * torch types.  Outputs are caller-allocated and the library reads no environment
 * variables; its only process-wide state is the mesh2sdf scratch pool (see
 * nglod_mesh2sdf / nglod_release_scratch).  Every function returns
 * 0 on success or a non-zero code (a cudaError_t, or one of the NGLOD_E*
 * values below for argument errors) -- it never prints and never aborts.  The
 * host side (nglod_b200/_lib.py -> lib/...) converts a non-zero return into a
 * Python RuntimeError.
 *
 * All device buffers are dense, row-major, fp32 unless stated; "borrowed"
 * means the callee never frees or keeps the pointer past the call's stream
 * work.  Outputs are caller-allocated (the torch-side shim allocates them the
 * way the reference extensions allocate their return tensors).
 *
 * Each entry point cites the reference interface it replaces
 * (paths relative to the nv-tlabs/nglod tree).
 */
#ifndef NGLOD_B200_H_
#define NGLOD_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NGLOD_ABI_VERSION 12
#define NGLOD_MAX_LODS 8

/* nglod_net_t.math_mode */
#define NGLOD_MATH_TC3XTF32 0 /* tcgen05 tensor cores, 3-pass split TF32 (FP32-level accuracy, |err| ~1e-6) */
#define NGLOD_MATH_FP32 1     /* CUDA cores, plain FP32 (|err| ~1e-7); the correctness anchor            */

/* argument-error codes (disjoint from cudaError_t values we can hit) */
#define NGLOD_EINVAL 10001   /* null pointer / negative size / bad lod        */
#define NGLOD_EUNSUPPORTED 10002 /* shape the kernels are not built for       */

/*
 * The OctreeSDF model, as the kernels see it.
 * Replaces: sdf-net/lib/models/OctreeSDF.py:38-88 (FeatureVolume grids + per-LOD
 * nn.Sequential(Linear(3+F,H), ReLU, Linear(H,1)) decoders).
 *
 * grids[i]: LOD i feature grid, CHANNELS-LAST: element (iz,iy,ix,c) at
 *           ((iz*(R+1)+iy)*(R+1)+ix)*feature_dim + c, R = grid_res[i]
 *           (this is the reference's fm[0,c,iz,iy,ix] held in
 *           torch.channels_last_3d memory format, so one corner's F channels
 *           are one contiguous, 16-byte-aligned run).
 * w0[i]:    [hidden_dim, in_dim] row-major (torch Linear.weight), where
 *           in_dim = feature_dim + (pos_invariant ? 0 : 3); xyz columns first.
 * b0[i]:    [hidden_dim];  w1[i]: [hidden_dim] (Linear(H,1).weight);  b1[i]: [1].
 * With a joint decoder all w0[i] (etc.) alias the same buffers.
 */
typedef struct nglod_net {
    int32_t num_lods;
    int32_t feature_dim;    /* 32 on the headline config                       */
    int32_t hidden_dim;     /* 128 on the headline config                      */
    int32_t pos_invariant;  /* 0: decoder input is [x,y,z,feat]; 1: [feat]     */
    int32_t math_mode;      /* NGLOD_MATH_* : how the 35->128 contraction runs  */
    int32_t reserved_;      /* keeps the pointer arrays 8-byte aligned          */
    int32_t grid_res[NGLOD_MAX_LODS];
    const float* grids[NGLOD_MAX_LODS];
    const float* w0[NGLOD_MAX_LODS];
    const float* b0[NGLOD_MAX_LODS];
    const float* w1[NGLOD_MAX_LODS];
    const float* b1[NGLOD_MAX_LODS];
    /* OPTIONAL inference accelerators, read by nglod_sdf_forward / nglod_sdf_finitediff / nglod_sphere_trace and -- for the
     * forward recompute of the single-grid backward, together with nglod_net_grad_t.summed -- by nglod_sdf_backward /
     * nglod_sdf_train_step (null = not provided: every kernel then gathers grids[]).  They are DERIVED data: whoever
     * writes grids[] must rebuild them (nglod_build_summed_grid) before the next call.
     * summed[i]:      the "prefix-summed" grid of LOD i, same layout and resolution as grids[i]:
     *                   summed[i][node] = sum_{l<=i} trilinear(grids[l], position of that node)
     *                 (nglod_build_summed_grid).  The LOD grids nest (grid_res[i] is a multiple of every coarser
     *                 grid_res[l]), so inside any cell of grid i every coarser interpolant is one trilinear polynomial
     *                 and trilinear interpolation reproduces it from its corner values: for EVERY x
     *                   trilinear(summed[i], x) == sum_{l<=i} trilinear(grids[l], x)
     *                 exactly in real arithmetic, to rounding in fp32.  The running sum of OctreeSDF.py:109-110 thus
     *                 costs ONE 8-corner gather instead of i+1 of them.
     * summed_fp16[i]: summed[i] as fp16 "x-pair lines" (nglod_pack_grid_fp16), used instead of summed[i] by the
     *                 tensor-core kernels when present.  The reference's own real-time format stores corner features
     *                 in fp16 too (SOL_NGLOD.py:73, `features.half()`). */
    const float* summed[NGLOD_MAX_LODS];
    const void* summed_fp16[NGLOD_MAX_LODS];
} nglod_net_t;

/* Gradient buffers shaped exactly like the parameters in nglod_net_t
 * (grids channels-last).  The kernels ACCUMULATE (+=) into them; the caller
 * zeroes them (autograd semantics).  Null entries are skipped. */
typedef struct nglod_net_grad {
    float* grids[NGLOD_MAX_LODS];
    float* w0[NGLOD_MAX_LODS];
    float* b0[NGLOD_MAX_LODS];
    float* w1[NGLOD_MAX_LODS];
    float* b1[NGLOD_MAX_LODS];
    /* OPTIONAL scratch for the single-grid backward: summed[i] is shaped like grids[i] and must be ALL ZERO on entry;
     * it is all zero again on return.  When net->summed[lod] and summed[0..lod] are given, nglod_sdf_backward /
     * nglod_sdf_train_step recompute the forward from the prefix-summed grid, scatter dL/d(summed[lod]) (8 corners per
     * query instead of 8*(lod+1)) into summed[lod] and then push it down the LOD chain with the transpose of the
     * prefix sum (dense restriction kernels, level by level -- the hat functions nest), adding into grids[0..lod]:
     * the same gradients as the per-LOD scatter, to fp32 rounding.
     * summed[i] MAY ALIAS grids[i] when grids[i] holds no other contribution on entry (the usual case: gradients zeroed
     * before the step): the level's gradient is then accumulated in place and the copy-out pass (140 MB of traffic
     * at the 65^3 level) is skipped; such a buffer is of course not zero on return. */
    float* summed[NGLOD_MAX_LODS];
    /* OPTIONAL scratch for the scatter into SMALL grids (single-grid path, grid_res <= 8): scatter_scratch_floats floats,
     * 16-byte aligned, ALL ZERO on entry and all zero again on return.  A grid of 125 or 729 nodes takes half a million
     * reductions per node and the L2 atomic units serialise per address (nglod_probe_scatter: 1.09 / 0.49 ms per 2^20
     * queries at grid_res 4 / 8 against 0.18 ms from grid_res 16 on); with this buffer the CTAs scatter into as many
     * private copies of the grid as fit (up to 64) and one small kernel folds them into summed[lod].  Null: one copy. */
    float* scatter_scratch;
    int64_t scatter_scratch_floats;
} nglod_net_grad_t;

/* ---- introspection ------------------------------------------------------ */
int nglod_abi_version(void);
/* static string: build arch, flags; never null */
const char* nglod_build_info(void);

/* Self-test of the tensor-core plumbing: D[128,128] = A[128,40] * B[128,40]^T (row-major fp32 device
 * buffers) through the same operand layout / descriptors / 3xTF32 sequence / TMEM read-back as the SDF kernels. */
int nglod_debug_tc_gemm(const float* A, const float* B, float* D, void* stream);

/* Measurement probe (bench.py's roofline denominator for the gather-bound kernels): reads, for n_queries pseudo-random
 * cells of a channels-last fp32 [(R+1)^3, 32] grid at `buf` (128-byte aligned), the 8 corner lines of the cell with
 * 8 lanes x 16 B per line -- the address stream of the SDF kernels' gather, no arithmetic (structured = 1) -- or
 * 8 * n_queries independent random 128-byte lines of the same buffer (structured = 0).  in_flight = 1..3: 8, 16 or 24
 * line loads in flight per lane.  Launch shape: 512-thread CTAs, as many per SM as fit next to smem_bytes of unused
 * dynamic shared memory each (the carve-out a real kernel would take from L1), at most ctas_per_sm (0 = no cap).
 * Bytes moved L2 -> SM: n_queries * 1024.  sink: device uint32[1], never written in practice.  No reference
 * counterpart (the reference has no measurement hooks). */
int nglod_probe_gather(const void* buf, int32_t grid_res, int64_t n_queries, int32_t in_flight,
                       int32_t structured, int32_t smem_bytes, int32_t ctas_per_sm, uint32_t seed,
                       uint32_t* sink, void* stream);
/* The scatter twin of nglod_probe_gather: adds a small value to the 8 corner lines of n_queries pseudo-random cells of
 * `buf` with 8 lanes x red.global.add.v4.f32 per line -- the address stream of the backward's grid-gradient scatter,
 * no arithmetic.  Bytes reduced into L2: n_queries * 1024.  Same launch-shape arguments, plus: ctas_per_sm < 0 launches
 * exactly -ctas_per_sm CTAs (fewer than one per SM: is the limit per SM or chip-wide?).  `buf` is modified. */
int nglod_probe_scatter(void* buf, int32_t grid_res, int64_t n_queries, int32_t smem_bytes, int32_t ctas_per_sm,
                        uint32_t seed, void* stream);

/* ---- ray vs unit cube ---------------------------------------------------
 * Replaces: f_aabb / aabb_kernel, sdf-net/lib/extensions/sol_nglod/
 * sol_nglod_kernel.cu:78-188 (exported to Python as sol_nglod.aabb, :190-192).
 * ray_o, ray_d: [n,3].  Outputs: x [n,3] (entry point, or ray_o when no hit),
 * t [n] (0 when no hit), hit [n] uint8 (0/1).  Rays whose origin is strictly
 * inside the cube report hit=0, x=ray_o, t=0 exactly like the reference.
 * Bit-exact with the reference kernel. */
int nglod_aabb(const float* ray_o, const float* ray_d, int64_t n,
               float* x, float* t, uint8_t* hit, void* stream);

/* ---- OctreeSDF.sdf(x, lod) forward --------------------------------------
 * Replaces: OctreeSDF.sdf, sdf-net/lib/models/OctreeSDF.py:94-155 with an
 * integer `lod` (F.grid_sample per LOD :46-57, running sum :109-110,
 * cat[x,feat] :115, louts[lod] :84-86).  x: [n,3] -> out: [n] (= [n,1]).
 * FP32 throughout (|err| vs the PyTorch path ~1e-7). */
int nglod_sdf_forward(const nglod_net_t* net, int32_t lod, const float* x,
                      int64_t n, float* out, void* stream);

/* Same model, all heads at once: out[l*n + i] = sdf(x_i, lod=l) for
 * l = 0..num_lods-1 (OctreeSDF.sdf(x, return_lst=True), OctreeSDF.py:152-153). */
int nglod_sdf_forward_all(const nglod_net_t* net, const float* x, int64_t n,
                          float* out, void* stream);

/* Summed interpolated features only (no decoder):
 * out[i, c] = sum_{l<=lod} trilinear(grids[l], x_i)[c],  out: [n, feature_dim].
 * Replaces: FeatureVolume.forward, sdf-net/lib/models/OctreeSDF.py:46-57 (one
 * LOD: pass a net whose only grid is that LOD) and the running sum :109-110.
 * Only grids / grid_res / num_lods / feature_dim of `net` are read. */
int nglod_sdf_features(const nglod_net_t* net, int32_t lod, const float* x,
                       int64_t n, float* out, void* stream);

/* summed[node] = sum_{l<=lod} trilinear(grids[l], node) for every node of grid `lod` (see nglod_net_t.summed).
 * dst: [(R+1)^3, feature_dim] fp32 channels-last, R = grid_res[lod], 16-byte aligned.  Node weights are exact
 * rationals ((ix mod k)/k with k = R / grid_res[l]).  NGLOD_EUNSUPPORTED if the grids do not nest.
 * If net->summed[lod-1] is set (the previous level, already built) the level is computed as its prolongation plus
 * grids[lod] -- build the levels in ascending order to get all of them for the price of the last. */
int nglod_build_summed_grid(const nglod_net_t* net, int32_t lod, float* dst, void* stream);

/* Build the half-precision gather layout of ONE grid: for every (z, y, x0 < R) a 128-byte line holding
 * {corner(x0) ch 0..31, corner(x0+1) ch 0..31} in fp16, so the two x-neighbours a trilinear sample needs are one
 * aligned line (one L1 wavefront) instead of two.  grid: channels-last fp32 [(R+1)^3, 32]; dst: (R+1)^2 * R * 128
 * bytes, 128-byte aligned. */
int nglod_pack_grid_fp16(const float* grid, int32_t grid_res, void* dst, void* stream);

/* ---- backward of sdf(x, lod) --------------------------------------------
 * Replaces: autograd of OctreeSDF.sdf driven from sdf-net/lib/trainer.py:339
 * (grid_sampler_3d_backward + Linear grads).  grad_out: [n] = dL/dd.
 * Accumulates dL/dgrids[i] (i<=lod) and dL/d{w0,b0,w1,b1}[lod] into `grad`;
 * heads != lod are untouched.  grad_x: [n,3] or null (dL/dx incl. PyTorch's
 * border-clip rule: the interpolation term is zeroed on an axis whose
 * unclipped index is <=0 or >=R). */
int nglod_sdf_backward(const nglod_net_t* net, int32_t lod, const float* x,
                       int64_t n, const float* grad_out,
                       const nglod_net_grad_t* grad, float* grad_x, void* stream);

/* ---- fused multi-LOD training step (forward + L2 loss + backward) --------
 * Replaces: Trainer.step_geometry's loss, sdf-net/lib/trainer.py:317-339:
 *   loss = sum_{l in lod_mask} sum_i (sdf(x_i,l) - gt_i)^2 * loss_scale
 * (the reference uses loss_scale = 1/batch).  Accumulates all gradients into
 * `grad`, adds the (scaled) loss into *loss_out (device, fp32; may be null).
 * lod_mask: bit l set <=> LOD l contributes; with NGLOD_LOSS_PER_LOD or'ed in, head l's loss goes to loss_out[l]
 * (an array of num_lods floats) instead of all heads summing into loss_out[0]. */
#define NGLOD_LOSS_PER_LOD 0x80000000u
int nglod_sdf_train_step(const nglod_net_t* net, uint32_t lod_mask,
                         const float* x, const float* gt, int64_t n,
                         float loss_scale, const nglod_net_grad_t* grad,
                         float* loss_out, void* stream);

/* ---- finite-difference gradient ------------------------------------------
 * Replaces: gradient(x, f, 'finitediff'), sdf-net/lib/diffutils.py:61-70:
 * out[i,k] = (sdf(x_i + h e_k) - sdf(x_i - h e_k)) / (2h), six evaluations,
 * NOT normalised.  x: [n,3] -> out: [n,3]. */
int nglod_sdf_finitediff(const nglod_net_t* net, int32_t lod, const float* x,
                         int64_t n, float h, float* out, void* stream);

/* ---- sphere tracer --------------------------------------------------------
 * Replaces: SphereTracer.forward, sdf-net/lib/tracer/SphereTracer.py:41-132
 * (aabb :53, march loop :74-115, box cull :119, finite-difference normals
 * :128-130) as ONE persistent kernel with the SDF evaluated inline.
 * Outputs: x [n,3], depth [n], hit [n] uint8, normal [n,3] (0 where !hit,
 * else grad / max(|grad|, 1e-5)).
 * queue: device int32[1] work counter, zeroed by the callee on `stream`.
 * stats: optional device uint64[2] {sdf evaluations, march iterations},
 *        accumulated (caller zeroes); may be null. */
typedef struct nglod_trace_opts {
    int32_t num_steps;       /* options.py:217  default 256                     */
    int32_t compute_normals; /* 1: finitediff normals on hits; 0: normals = 0   */
    /* doubles on purpose: these are Python floats in the reference; the callee
     * rounds them to fp32 the way torch does when a tensor meets a Python
     * scalar (e.g. the oscillation threshold is float32(min_dis*3.0)). */
    double step_size;        /* options.py:219  default 1.0                     */
    double min_dis;          /* options.py:221  default 3e-4                    */
    double far;              /* camera_clamp[1], options.py:209 default 10      */
    double normal_h;         /* diffutils.py:62  1/(64*3)                       */
    /* 0: the persistent dense tracer takes every SM (one CTA each: a CTA holds the SM's whole register file and shared
     * memory, so no kernel of another stream runs beside it).  > 0: at most this many CTAs -- leaves SMs to concurrent
     * work, e.g. the shading of the previous ray range (Renderer.shade_images).  Results do not depend on it. */
    int32_t max_ctas;
    int32_t reserved_;       /* 0 */
} nglod_trace_opts_t;

int nglod_sphere_trace(const nglod_net_t* net, int32_t lod,
                       const float* ray_o, const float* ray_d, int64_t n,
                       const nglod_trace_opts_t* opts,
                       float* x, float* depth, uint8_t* hit, float* normal,
                       int32_t* queue, unsigned long long* stats, void* stream);
/* The same frame as packed per-ray records, each written with 16-byte stores when the ray retires.  Meant for a PINNED
 * HOST buffer (cudaHostAlloc memory is device-addressable under unified addressing): the results cross PCIe as posted
 * writes while the kernel runs, so a host caller needs no device->host copy of the frame after the trace and no
 * chunking -- SphereTracer.trace_lookat_host(packed=True).
 *   hit == NULL: packed is [n][8] fp32 (32-byte aligned) = {depth, nx, ny, nz, hit (uint32 0 / 1), x, y, z};
 *   hit != NULL: packed is [n][4] fp32 (16-byte aligned) = {depth, nx, ny, nz} and the hit flags go to hit [n] (bytes,
 *                normally DEVICE memory: n bytes to copy afterwards instead of 17 n).
 * No reference counterpart (the reference's `.cpu()` copies every RenderBuffer field after the frame). */
int nglod_sphere_trace_packed(const nglod_net_t* net, int32_t lod, const float* ray_o, const float* ray_d,
                              int64_t n, const nglod_trace_opts_t* opts, float* packed, uint8_t* hit, int32_t* queue,
                              unsigned long long* stats, void* stream);

/* One camera frame, host to host, in ONE call: window coordinates (width + height floats, pinned host or device memory)
 * -> nglod_generate_rays -> nglod_sphere_trace_packed -> (hit_copy != NULL) the n hit bytes copied to hit_copy, all
 * queued on `stream`; the caller synchronises the stream and reads `packed` / `hit_copy`.  workspace: device memory,
 * (6 n + width + height) floats (the rays and the window), n = width * height.  packed / hit as in
 * nglod_sphere_trace_packed.  Replaces, for a host caller, Renderer.render_lookat + `.cpu()`
 * (sdf-net/lib/renderer.py:89-107: look_at -> tracer -> RenderBuffer fields copied back one by one). */
int nglod_sphere_trace_camera(const nglod_net_t* net, int32_t lod, const float* origin, const float* view,
                              const float* right, const float* up, float tan_half_fov, int32_t ortho,
                              const float* window_x, const float* window_y, int32_t width, int32_t height,
                              const nglod_trace_opts_t* opts, float* workspace, float* packed, uint8_t* hit,
                              uint8_t* hit_copy, int32_t* queue, unsigned long long* stats, void* stream);

/* ---- mesh -> signed distance ----------------------------------------------
 * Replaces: mesh2sdf_gpu_fast_nopre -> kernel_mesh2sdf_quad + kernel_quad_aggr,
 * sdf-net/lib/extensions/mesh2sdf_cuda/mesh2sdf_kernel.cu:307-616,895-927
 * (exported as mesh2sdf.mesh2sdf_gpu, :1007-1012).
 * points: [n,3]; tris: [T,3,3] (= V[F]); dist: [n] signed distance
 * (negative iff all 13 stab directions see a triangle on both sides).
 * No [64,n,26] temporaries.  Small batches walk all n x T pairs with the triangles staged through shared memory;
 * batches >= 16 k points take an output-sensitive path (nearest triangle through a sphere hierarchy over
 * Morton-sorted triangles, sign through 13 projected point grids) whose results are bit-identical to the walk.
 * Scratch (records, bins: ~100 B per point + 400 B per triangle) is stream-ordered, from a memory pool the
 * library keeps per device -- the one piece of process-wide state behind this ABI (a cudaMemPool_t per device,
 * created on first use, mutex-guarded, never trimmed by synchronisation); nglod_release_scratch() hands its
 * unused memory back to the driver.  No host synchronisation.
 * nglod_mesh2sdf_ex: flags = NGLOD_M2S_FORCE_WALK makes every batch take the brute-force walk (the A/B reference of
 * the large-batch path; results are bit-identical either way). */
#define NGLOD_M2S_FORCE_WALK 1u
int nglod_mesh2sdf(const float* points, int64_t n, const float* tris,
                   int64_t num_tris, float* dist, void* stream);
int nglod_mesh2sdf_ex(const float* points, int64_t n, const float* tris,
                      int64_t num_tris, float* dist, uint32_t flags, void* stream);
/* Trim the current device's mesh2sdf scratch pool to zero retained bytes (work still in flight keeps its memory). */
int nglod_release_scratch(void);

/* ---- training-point sampler -------------------------------------------------
 * Replaces the host-side torch samplers of sdf-net/lib/torchgp/: area_weighted_distribution.py:26-45,
 * random_face.py:27-47, sample_surface.py:27-52, sample_near_surface.py:27-45, sample_uniform.py:25-31 and
 * their driver point_sample.py:29-57 (called from MeshDataset.resample, MeshDataset.py:71-93).
 * nglod_mesh_area_cdf: V [#V,3] fp32, F [num_faces,3] int64 -> cdf [num_faces] fp32, the inclusive cumulative
 *   face areas (non-decreasing; accumulated in double), built once per mesh.
 * nglod_sample_mesh: `techniques` is a HOST array of num_techniques (<= 32) codes; writes
 *   pts [num_techniques * samples_per_technique, 3], technique-major in the order given (point_sample's
 *   concatenation); face_idx (nullable) [same count] int32 = the face a surface sample lies on (-1 for RAND),
 *   from which the caller gathers per-face normals (sample_surface returns them).
 *   RAND: U[-1,1)^3.  TRACE: face ~ area, u = sqrt(r1), v = r2, p = (1-u) a + u (1-v) b + u v c.
 *   NEAR: TRACE + N(0,1) * variance per coordinate (the reference multiplies by `variance`, default 0.01).
 *   Randoms come from Philox-4x32-10 keyed by `seed`, counter = sample index: the same seed reproduces the same
 *   points; streams differ from torch's, so parity with the reference is distributional.
 *   V / F / cdf may be NULL when every technique is RAND. */
#define NGLOD_SAMPLE_RAND 0
#define NGLOD_SAMPLE_NEAR 1
#define NGLOD_SAMPLE_TRACE 2
int nglod_mesh_area_cdf(const float* V, const int64_t* F, int64_t num_faces, float* cdf, void* stream);
int nglod_sample_mesh(const float* V, const int64_t* F, int64_t num_faces, const float* cdf,
                      const int* techniques, int num_techniques, int64_t samples_per_technique,
                      float variance, uint64_t seed, float* pts, int* face_idx, void* stream);

/* ---- sparse-octree (SPC) ray traversal -------------------------------------
 * Replaces: spc_raytrace, sol-renderer/include/spc/spc/spc_raytrace_cuda.cpp:125-199 + kernels
 * spc_raytrace_cuda_kernel.cu:51-265 (= Kaolin's unbatched_raytrace, sdf-net/app/spc/SPC.py:104-107).
 * octree: child-mask bytes, breadth first; prefix: exclusive sum of their popcounts (int32, same length);
 * points: [psize,4] int16 voxel coordinates of every level in Morton order; pyramid_sum: HOST int32[level+2],
 * first point of each level; level: depth of the octree; target_level <= level.
 * Output "nuggets" [total,2] int32 = (ray index, point index within target_level), sorted by ray, each ray's
 * run front to back -- identical, nugget for nugget, to the reference's level-synchronous traversal.
 * Two calls: _count writes offsets[n+1] (exclusive scan of per-ray nugget counts; offsets[n] = total) using
 * scan_ws (int32[n/1024+2]); the caller reads the total, allocates, and _fill writes the nuggets. */
int nglod_spc_raytrace_count(const uint8_t* octree, const int32_t* prefix, const int16_t* points,
                             const int32_t* pyramid_sum, int32_t level, int32_t target_level,
                             const float* ray_o, const float* ray_d, int64_t n,
                             int32_t* offsets, int32_t* scan_ws, void* stream);
int nglod_spc_raytrace_fill(const uint8_t* octree, const int32_t* prefix, const int16_t* points,
                            const int32_t* pyramid_sum, int32_t level, int32_t target_level,
                            const float* ray_o, const float* ray_d, int64_t n,
                            const int32_t* offsets, int32_t* nuggets, void* stream);

/* info[i] = 1 where nugget i starts a new ray's run.  Replaces: d_MarkUniqueRays, sol-renderer/sdfRenderer.cu:108-120
 * (= Kaolin's mark_first_hit, sdf-net/app/spc/SPCTracer.py:56). */
int nglod_spc_mark_first_hit(const int32_t* nuggets, int64_t m, int32_t* info, void* stream);

/* First voxel of each ray's run that contains `query` (no advance) or that the ray enters (t += d, x = o + t*d);
 * rays whose run is exhausted get cond = 0, t = 100.  Rays without nuggets (and rays with active[ray] == 0 when
 * `active` is non-null) are left untouched.  level_points: the [*,4] int16 points of the nugget level.
 * Replaces: ray_aabb_kernel / ray_aabb, sol-renderer/include/solr/solr/gfx/ray_aabb.cuh:42-192
 * (= Kaolin's unbatched_ray_aabb, SPCTracer.py:62,99). */
int nglod_spc_ray_aabb(const int32_t* nuggets, const int32_t* offsets, int64_t n_rays,
                       const int16_t* level_points, int32_t level, const float* ray_o, const float* ray_d,
                       const float* query, const uint8_t* active, float* x, float* t, uint8_t* cond,
                       int32_t* pidx, void* stream);

/* ---- sparse OctreeSDF: in-voxel evaluation and tracing ---------------------------
 * The model of the reference's real-time renderer (sol-renderer/SDF.cu:65-216): corner features + per-voxel
 * 8-corner "trinkets" with a parent link, one 35->128->1 decoder per LOD.  All LODs concatenated, coarse first. */
typedef struct nglod_sparse_net {
    int32_t num_lods;
    int32_t base_lod;        /* LOD l lives on octree level l + base_lod                          */
    int32_t feature_dim;     /* 32                                                                 */
    int32_t hidden_dim;      /* 128                                                                */
    int32_t math_mode;       /* NGLOD_MATH_*                                                       */
    int32_t pos_invariant;   /* 1: decoders take the features only (the reference's NeuralSPC, app/spc/NeuralSPC.py:91-95);
                                0: [x, y, z, features] like OctreeSDF                                  */
    int32_t lod_voxel_offset[NGLOD_MAX_LODS + 2]; /* first voxel row of each LOD (and one past the last)  */
    const float* corner_feats;   /* [NC, feature_dim]                                              */
    const int32_t* trinkets;     /* [NV, 8] rows of corner_feats; corner k = bx + 2*by + 4*bz      */
    const int32_t* parents;      /* [NV] voxel row one LOD up, -1 at LOD 0                         */
    const int16_t* voxels;       /* [NV, 4] integer voxel coordinates at the voxel's own level     */
    const float* w0[NGLOD_MAX_LODS];
    const float* b0[NGLOD_MAX_LODS];
    const float* w1[NGLOD_MAX_LODS];
    const float* b1[NGLOD_MAX_LODS];
    /* OPTIONAL [NC, feature_dim], same rows as corner_feats: the row of a LOD-l corner holds
     *   sum_{k<=l} trilinear(corner features of the level-k ancestor voxel, position of that corner)
     * (the sparse twin of nglod_net_t.summed: octree levels nest, so ONE 8-corner sample of the requested LOD's voxel
     * equals the sum over its parent chain).  When given, the kernels read it and never touch `parents`. */
    const float* corner_feats_summed;
} nglod_sparse_net_t;

/* out[i] = decoder_lod([x_i, sum_{l<=lod} trilinear(corner features of the voxel chain of pidx_i)]).
 * pidx: voxel index within LOD `lod` (level-local, as in the nuggets).  Weights are NOT clamped: a point slightly
 * outside its voxel extrapolates, as in the reference.
 * Replaces: sparse_grid_sample_kernel + 2x torch::addmm, sol-renderer/include/solr/solr/sdf/sparse_grid_sample.cuh:31-109,
 * SDF.cu:412-413 (Python twin: NeuralSPC.sdf / SPC.interpolate, sdf-net/app/spc/NeuralSPC.py:104-145, SPC.py:92-102). */
int nglod_sparse_sdf_forward(const nglod_sparse_net_t* net, int32_t lod, const float* x, const int32_t* pidx,
                             int64_t n, float* out, void* stream);

/* Backward of nglod_sparse_sdf_forward for training a natively sparse model (the reference's app/spc: NeuralSPC.sdf under
 * autograd, NeuralSPC.py:104-145).  grad_out: [n] = dL/dd.  ACCUMULATES dL/d(corner_feats) into grad_corner_feats
 * ([NC, feature_dim], every level of the query's parent chain) and dL/d{w0,b0,w1,b1}[lod] into gw0..gb1 (null = skip).
 * Always walks the parent chain (corner_feats_summed is an inference-only accelerator). */
int nglod_sparse_sdf_backward(const nglod_sparse_net_t* net, int32_t lod, const float* x, const int32_t* pidx, int64_t n,
                              const float* grad_out, float* grad_corner_feats, float* gw0, float* gb0, float* gw1,
                              float* gb1, void* stream);

/* Fused training step of one head of a natively sparse model: forward of LOD `lod`, L = loss_scale * sum_i (d_i - gt_i)^2
 * ADDED to *loss_out (nullable), and the backward of nglod_sparse_sdf_backward with dL/dd = 2 loss_scale (d - gt), in ONE
 * kernel - the sparse twin of nglod_sdf_train_step.  Replaces, per head, NeuralSPC.sdf + the L2 loss + autograd of the
 * reference's sparse trainer (sdf-net/app/spc/NeuralSPC.py:104-145, the loss of lib/trainer.py:317-327). */
int nglod_sparse_sdf_train_step(const nglod_sparse_net_t* net, int32_t lod, const float* x, const int32_t* pidx,
                                const float* gt, int64_t n, float loss_scale, float* grad_corner_feats,
                                float* gw0, float* gb0, float* gw1, float* gb1, float* loss_out, void* stream);

/* In-voxel sphere tracing over the nuggets of nglod_spc_raytrace (level lod + base_lod), ONE persistent kernel:
 * first voxel (ray_aabb) -> [sparse sdf -> step -> re-locate the voxel from the new position] x num_steps -> central-
 * difference normals (h = normal_h, same voxel) on hits.  hit = |d| < min_dis or |d + dprev|/2 < 5*min_dis;
 * a ray stays alive while t < far and it still finds a voxel; rays that run out of voxels get depth 100.
 * Replaces: SDF::sphereTrace / getNormal, sol-renderer/SDF.cu:218-472 with step.cuh:31-86 and ray_aabb.cuh:104-192
 * (Python twin: SPCTracer.forward, sdf-net/app/spc/SPCTracer.py:44-113).  opts->step_size is ignored (the reference
 * renderer has none); pidx_out (optional) receives each ray's last voxel. */
int nglod_spc_sphere_trace(const nglod_sparse_net_t* net, int32_t lod, const int32_t* nuggets, const int32_t* offsets,
                           const float* ray_o, const float* ray_d, int64_t n, const nglod_trace_opts_t* opts,
                           float* x, float* depth, uint8_t* hit, float* normal, int32_t* pidx_out,
                           int32_t* queue, unsigned long long* stats, void* stream);
/* The frame's own fast path (no reference counterpart: the reference renderer takes the two-pass nugget list above).
 * nglod_spc_raytrace_runs: ONE pass over the rays -- walk, reserve the ray's run with a warp-aggregated atomicAdd, write
 * it.  nuggets [capacity, 2] holds the runs in arbitrary ray order, each run in the reference's front-to-back order:
 * the run of ray i is nuggets[run_begin[i] .. run_end[i]).  cursor: device int32[2], set by the call: cursor[0] = slots
 * reserved, cursor[1] = 1 if capacity was too small (some runs were dropped: take nglod_spc_raytrace_count / _fill).
 * No scan, no second traversal launch, no host read between traversal and tracing.
 * nglod_spc_sphere_trace_runs: nglod_spc_sphere_trace over such runs (nglod_spc_sphere_trace(.., offsets, ..) is this
 * call with run_begin = offsets, run_end = offsets + 1). */
int nglod_spc_raytrace_runs(const uint8_t* octree, const int32_t* prefix, const int16_t* points,
                            const int32_t* pyramid_sum, int32_t level, int32_t target_level,
                            const float* ray_o, const float* ray_d, int64_t n, int64_t capacity,
                            int32_t* nuggets, int32_t* run_begin, int32_t* run_end, int32_t* cursor, void* stream);
int nglod_spc_sphere_trace_runs(const nglod_sparse_net_t* net, int32_t lod, const int32_t* nuggets,
                                const int32_t* run_begin, const int32_t* run_end, const float* ray_o,
                                const float* ray_d, int64_t n, const nglod_trace_opts_t* opts, float* x,
                                float* depth, uint8_t* hit, float* normal, int32_t* pidx_out, int32_t* queue,
                                unsigned long long* stats, void* stream);

/* Host-only (no device work): the camera frame of look_at -- basis [4][3] = origin, view, right, up from the eye `from`
 * and the target `to` (3 floats each), in the float32 arithmetic torch's CPU kernels use, so that it equals
 *   view = F.normalize(to - from); right = F.normalize(cross(view, (0,1,0))); up = F.normalize(cross(right, view))
 * (sdf-net/lib/geoutils.py:180-188) bit for bit.  Feeds nglod_generate_rays / nglod_sphere_trace_camera. */
int nglod_camera_basis(const float* from, const float* to, float* basis);

/* ---- renderer entry-point helpers ------------------------------------------------
 * Camera rays, x-major (ray = ix*height + iy).  origin/view/right/up: HOST float[3] (already normalised, as computed
 * by look_at, sdf-net/lib/geoutils.py:180-188); window_x [width] / window_y [height]: DEVICE window coordinates
 * including the per-column / per-row jitter of normalized_grid (geoutils.py:140-154).  ortho: 0 = perspective.
 * Replaces the tensor expression of geoutils.py:189-204. */
int nglod_generate_rays(const float* origin, const float* view, const float* right, const float* up,
                        float tan_half_fov, int32_t ortho, const float* window_x, const float* window_y,
                        int32_t width, int32_t height, float* ray_o, float* ray_d, void* stream);

/* Matcap shading of a traced frame on the device: uv = spherical_envmap(view, normal) (geoutils.py:253-275),
 * rgb = bilinear(matcap[nu,nv,nc], uv)/255 on hits; misses get rgb = 1 and normal = 1 (renderer.py:279-296, which
 * does this lookup on the HOST through scipy).  view/normal/rgb: [n,3]; hit: [n] u8; matcap indexed [u][v][c]. */
int nglod_shade_matcap(const float* view, float* normal, const uint8_t* hit, const float* matcap,
                       int32_t nu, int32_t nv, int32_t nc, int64_t n, float* rgb, void* stream);

/* ---- Adam on a flat fp32 parameter buffer ----------------------------------
 * Replaces: torch.optim.Adam(lr) as set up by Trainer.set_optimizer,
 * sdf-net/lib/trainer.py:178-189 (betas .9/.999, eps 1e-8, no weight decay).
 * bias corrections are passed in by the host: bc1 = 1-beta1^t, bc2 = 1-beta2^t. */
int nglod_adam_step(float* param, const float* grad, float* exp_avg,
                    float* exp_avg_sq, int64_t n, float lr, float beta1,
                    float beta2, float eps, float bc1, float bc2, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NGLOD_B200_H_ */
