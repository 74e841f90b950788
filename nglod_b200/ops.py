"""Torch-facing wrappers over the C ABI (include/nglod_b200.h).

PyTorch here is plumbing only: it owns device memory and the stream; every op
below hands raw pointers to a hand-written sm_100a kernel.  Outputs are
allocated the way the reference extensions allocate theirs (new tensors on the
inputs' device; sol_nglod_kernel.cu:167-173, mesh2sdf_kernel.cu:901-906).
"""
import ctypes

import torch

from . import _lib
from ._lib import NetStruct, NetGradStruct, TraceOpts, MAX_LODS


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _f32c(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor -- nglod_b200 has no CPU path")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class NetView:
    """A borrowed view of an OctreeSDF's parameters as an nglod_net_t.

    grids: list of [1, F, R+1, R+1, R+1] tensors in channels_last_3d memory
    format (i.e. physically [z, y, x, F]); decoders: list of (w0, b0, w1, b1).
    Keeps the tensors alive for as long as the view is.
    """

    def __init__(self, grids, decoders, pos_invariant=False, math_mode=0, summed=None, summed_half=None):
        self.grids = grids
        self.decoders = decoders
        self.summed, self.summed_half = summed, summed_half      # optional inference accelerators (kept alive here)
        n = len(grids)
        if n < 1 or n > MAX_LODS:
            raise RuntimeError(f"num_lods must be in [1, {MAX_LODS}]")
        s = NetStruct()
        s.num_lods = n
        s.feature_dim = grids[0].shape[1]
        s.hidden_dim = decoders[0][0].shape[0]
        s.pos_invariant = 1 if pos_invariant else 0
        s.math_mode = int(math_mode)
        for i, g in enumerate(grids):
            if not g.is_cuda:
                raise RuntimeError("OctreeSDF parameters must live on a CUDA device (no CPU path)")
            if g.dtype != torch.float32 or not g.is_contiguous(memory_format=torch.channels_last_3d):
                raise RuntimeError("feature grids must be fp32 in torch.channels_last_3d memory format")
            s.grid_res[i] = g.shape[-1] - 1
            s.grids[i] = g.data_ptr()
        for i in range(n):
            w0, b0, w1, b1 = decoders[i if len(decoders) > 1 else 0]
            for t in (w0, b0, w1, b1):
                if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                    raise RuntimeError("decoder parameters must be contiguous fp32 CUDA tensors")
            s.w0[i], s.b0[i], s.w1[i], s.b1[i] = w0.data_ptr(), b0.data_ptr(), w1.data_ptr(), b1.data_ptr()
        for i, sg in enumerate(summed or []):
            if sg is None:
                continue
            if sg.dtype != torch.float32 or not sg.is_cuda or sg.shape != grids[i].shape or \
                    not sg.is_contiguous(memory_format=torch.channels_last_3d):
                raise RuntimeError("summed[i] must look like grids[i] (fp32, channels_last_3d) -- see build_summed_grid")
            s.summed[i] = sg.data_ptr()
        for i, h in enumerate(summed_half or []):
            if h is None:
                continue
            R = grids[i].shape[-1] - 1
            if h.dtype != torch.uint8 or not h.is_cuda or h.numel() != (R + 1) * (R + 1) * R * 128 or h.data_ptr() % 128:
                raise RuntimeError("summed_half[i] must be the 128-byte aligned uint8 buffer pack_grid_fp16 produced")
            if not summed or summed[i] is None:
                raise RuntimeError("summed_half[i] needs summed[i] (the CUDA-core kernels read the fp32 copy)")
            s.summed_fp16[i] = h.data_ptr()
        self.struct = s
        self.device = grids[0].device

    @classmethod
    def grids_only(cls, grids):
        """View that carries only feature grids (enough for nglod_sdf_features)."""
        self = cls.__new__(cls)
        s = NetStruct()
        s.num_lods = len(grids)
        s.feature_dim = grids[0].shape[1]
        for i, g in enumerate(grids):
            if not g.is_cuda or g.dtype != torch.float32 or not g.is_contiguous(memory_format=torch.channels_last_3d):
                raise RuntimeError("feature grids must be fp32 CUDA tensors in channels_last_3d memory format")
            s.grid_res[i] = g.shape[-1] - 1
            s.grids[i] = g.data_ptr()
        self.struct, self.grids, self.decoders, self.device = s, list(grids), [], grids[0].device
        return self

    @property
    def num_lods(self):
        return self.struct.num_lods


def make_grad_struct(view, grid_grads, dec_grads, summed_scratch=None, scatter_scratch=None):
    """grid_grads[i] / dec_grads[i] = (gw0, gb0, gw1, gb1) or None entries.  summed_scratch[i]: all-zero buffers shaped
    like the grids (nglod_net_grad_t.summed) that enable the single-grid backward; zero again when the call returns.
    scatter_scratch: an all-zero flat fp32 buffer (nglod_net_grad_t.scatter_scratch) for private scatter copies of the
    small grids; zero again when the call returns."""
    g = NetGradStruct()
    if scatter_scratch is not None:
        if scatter_scratch.dtype != torch.float32 or not scatter_scratch.is_cuda or not scatter_scratch.is_contiguous():
            raise RuntimeError("scatter_scratch must be a contiguous fp32 CUDA tensor (zero-filled)")
        g.scatter_scratch = scatter_scratch.data_ptr()
        g.scatter_scratch_floats = scatter_scratch.numel()
    for i, t in enumerate(summed_scratch or []):
        if t is not None:
            if t.shape != view.grids[i].shape or t.dtype != torch.float32 or not t.is_cuda or \
                    not t.is_contiguous(memory_format=torch.channels_last_3d):
                raise RuntimeError("summed_scratch[i] must look like grids[i] (fp32, channels_last_3d, zero-filled)")
            g.summed[i] = t.data_ptr()
    for i in range(view.num_lods):
        gg = grid_grads[i] if grid_grads is not None else None
        g.grids[i] = gg.data_ptr() if gg is not None else 0
        dg = dec_grads[i] if dec_grads is not None else None
        if dg is not None:
            g.w0[i], g.b0[i], g.w1[i], g.b1[i] = (t.data_ptr() if t is not None else 0 for t in dg)
    return g


# --------------------------------------------------------------------------- aabb
def aabb(ray_o, ray_d):
    """Drop-in for `sol_nglod.aabb` (sol_nglod_kernel.cu:161-192): returns (x [N,3], t [N,1], hit [N] bool)."""
    lib = _lib.load()
    ray_o = _f32c(ray_o, "ray_o")
    ray_d = _f32c(ray_d, "ray_d")
    n = ray_o.shape[0]
    x = torch.empty_like(ray_o)
    t = torch.empty(n, 1, device=ray_o.device, dtype=torch.float32)
    hit = torch.empty(n, device=ray_o.device, dtype=torch.bool)
    with torch.cuda.device(ray_o.device):
        _lib.check(lib.nglod_aabb(_ptr(ray_o), _ptr(ray_d), n, _ptr(x), _ptr(t), _ptr(hit), _stream()), "nglod_aabb")
    return x, t, hit


# --------------------------------------------------------------------------- sdf
def sdf_forward(view, lod, x):
    lib = _lib.load()
    x = _f32c(x, "x")
    n = x.shape[0]
    out = torch.empty(n, 1, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(lib.nglod_sdf_forward(ctypes.byref(view.struct), lod, _ptr(x), n, _ptr(out), _stream()),
                   "nglod_sdf_forward")
    return out


def build_summed_grid(view, lod, out=None):
    """Prefix-summed grid of LOD `lod` (nglod_net_t.summed, nglod_build_summed_grid): a tensor shaped and laid out like
    view.grids[lod] whose single trilinear sample equals the sum of the samples of grids 0..lod."""
    lib = _lib.load()
    g = view.grids[lod]
    if out is None:
        out = torch.empty_like(g, memory_format=torch.preserve_format)
    with torch.cuda.device(g.device):
        _lib.check(lib.nglod_build_summed_grid(ctypes.byref(view.struct), lod, _ptr(out), _stream()),
                   "nglod_build_summed_grid")
    return out


def pack_grid_fp16(fm, out=None):
    """fp16 "x-pair line" copy of one feature grid for the inference kernels (include/nglod_b200.h,
    nglod_pack_grid_fp16): [1,F,S,S,S] channels_last_3d fp32 -> uint8 [S*S*(S-1)*128]."""
    lib = _lib.load()
    if not fm.is_cuda or fm.dtype != torch.float32 or not fm.is_contiguous(memory_format=torch.channels_last_3d):
        raise RuntimeError("feature grids must be fp32 CUDA tensors in channels_last_3d memory format")
    S = fm.shape[-1]
    nbytes = S * S * (S - 1) * 128
    if out is None:
        out = torch.empty(nbytes, device=fm.device, dtype=torch.uint8)   # torch's allocator aligns to 512 B
    with torch.cuda.device(fm.device):
        _lib.check(lib.nglod_pack_grid_fp16(_ptr(fm), S - 1, _ptr(out), _stream()), "nglod_pack_grid_fp16")
    return out


def sdf_forward_all(view, x):
    lib = _lib.load()
    x = _f32c(x, "x")
    n = x.shape[0]
    out = torch.empty(view.num_lods, n, 1, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(lib.nglod_sdf_forward_all(ctypes.byref(view.struct), _ptr(x), n, _ptr(out), _stream()),
                   "nglod_sdf_forward_all")
    return out


def sdf_features(view, lod, x):
    lib = _lib.load()
    x = _f32c(x, "x")
    n = x.shape[0]
    out = torch.empty(n, view.struct.feature_dim, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(lib.nglod_sdf_features(ctypes.byref(view.struct), lod, _ptr(x), n, _ptr(out), _stream()),
                   "nglod_sdf_features")
    return out


def sdf_backward(view, lod, x, grad_out, grid_grads, dec_grad, want_grad_x=False, summed_scratch=None, scatter_scratch=None):
    """Accumulate into grid_grads[0..lod] (channels_last_3d, like the params) and dec_grad=(gw0,gb0,gw1,gb1)."""
    lib = _lib.load()
    x = _f32c(x, "x")
    grad_out = _f32c(grad_out, "grad_out").reshape(-1)
    n = x.shape[0]
    dec = [None] * view.num_lods
    dec[lod] = dec_grad
    gs = make_grad_struct(view, grid_grads, dec, summed_scratch, scatter_scratch)
    gx = torch.empty_like(x) if want_grad_x else None
    with torch.cuda.device(x.device):
        _lib.check(lib.nglod_sdf_backward(ctypes.byref(view.struct), lod, _ptr(x), n, _ptr(grad_out),
                                          ctypes.byref(gs), _ptr(gx), _stream()), "nglod_sdf_backward")
    return gx


def sdf_train_step(view, lod_mask, x, gt, loss_scale, grid_grads, dec_grads, loss_out=None, summed_scratch=None,
                   scatter_scratch=None):
    lib = _lib.load()
    x = _f32c(x, "x")
    gt = _f32c(gt, "gt").reshape(-1)
    n = x.shape[0]
    gs = make_grad_struct(view, grid_grads, dec_grads, summed_scratch, scatter_scratch)
    with torch.cuda.device(x.device):
        _lib.check(lib.nglod_sdf_train_step(ctypes.byref(view.struct), lod_mask, _ptr(x), _ptr(gt), n,
                                            float(loss_scale), ctypes.byref(gs), _ptr(loss_out), _stream()),
                   "nglod_sdf_train_step")


def sdf_finitediff(view, lod, x, h=1.0 / (64.0 * 3.0)):
    lib = _lib.load()
    x = _f32c(x, "x")
    n = x.shape[0]
    out = torch.empty(n, 3, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(lib.nglod_sdf_finitediff(ctypes.byref(view.struct), lod, _ptr(x), n, float(h), _ptr(out),
                                            _stream()), "nglod_sdf_finitediff")
    return out


# --------------------------------------------------------------------------- tracer
def sphere_trace(view, lod, ray_o, ray_d, num_steps=256, step_size=1.0, min_dis=0.0003, far=10.0,
                 normal_h=1.0 / (64.0 * 3.0), compute_normals=True, stats=None, out=None, queue=None, max_ctas=0):
    """One persistent kernel: returns x [N,3], depth [N,1], hit [N] bool, normal [N,3].
    out=(x, depth, hit, normal): caller-provided contiguous device buffers (e.g. slices of a frame buffer).
    max_ctas > 0: at most that many CTAs (one per SM) -- leaves SMs to kernels of other streams; same results."""
    lib = _lib.load()
    ray_o = _f32c(ray_o, "ray_o")
    ray_d = _f32c(ray_d, "ray_d")
    n = ray_o.shape[0]
    dev = ray_o.device
    if out is None:
        x = torch.empty(n, 3, device=dev, dtype=torch.float32)
        depth = torch.empty(n, 1, device=dev, dtype=torch.float32)
        hit = torch.empty(n, device=dev, dtype=torch.bool)
        normal = torch.empty(n, 3, device=dev, dtype=torch.float32)
    else:
        x, depth, hit, normal = out
        for t, shape, dt in ((x, (n, 3), torch.float32), (depth, (n, 1), torch.float32), (hit, (n,), torch.bool),
                             (normal, (n, 3), torch.float32)):
            if tuple(t.shape) != shape or t.dtype != dt or t.device != dev or not t.is_contiguous():
                raise RuntimeError("sphere_trace: out buffers must be contiguous device tensors x[N,3] depth[N,1] hit[N] normal[N,3]")
    if queue is None:
        queue = torch.empty(1, device=dev, dtype=torch.int32)
    opts = TraceOpts(int(num_steps), 1 if compute_normals else 0, float(step_size), float(min_dis), float(far),
                     float(normal_h), int(max_ctas), 0)
    with torch.cuda.device(dev):
        _lib.check(lib.nglod_sphere_trace(ctypes.byref(view.struct), lod, _ptr(ray_o), _ptr(ray_d), n,
                                          ctypes.byref(opts), _ptr(x), _ptr(depth), _ptr(hit), _ptr(normal),
                                          _ptr(queue), _ptr(stats), _stream()), "nglod_sphere_trace")
    return x, depth, hit, normal


def sphere_trace_packed(view, lod, ray_o, ray_d, packed, hit=None, num_steps=256, step_size=1.0, min_dis=0.0003, far=10.0,
                        normal_h=1.0 / (64.0 * 3.0), compute_normals=True, stats=None, queue=None):
    """The frame as packed per-ray records (nglod_sphere_trace_packed).  `packed` may be a device tensor or a PINNED HOST
    tensor: host memory is written by the kernel itself (posted PCIe writes as rays retire), so a host caller needs no
    copy of the frame after the trace.
      hit is None : packed is [N, 8] fp32, 32-byte records {depth, nx, ny, nz, hit (u32), x, y, z};
      hit [N] bool / uint8 on the rays' device: packed is [N, 4] fp32, 16-byte records {depth, nx, ny, nz}."""
    lib = _lib.load()
    ray_o = _f32c(ray_o, "ray_o")
    ray_d = _f32c(ray_d, "ray_d")
    n, dev = ray_o.shape[0], ray_o.device
    width = 8 if hit is None else 4
    if tuple(packed.shape) != (n, width) or packed.dtype != torch.float32 or not packed.is_contiguous() or \
            not (packed.device == dev or (packed.device.type == "cpu" and packed.is_pinned())):
        raise RuntimeError(f"sphere_trace_packed: packed must be a contiguous fp32 [N, {width}] tensor on the rays' device or in "
                           "pinned host memory")
    if hit is not None and (tuple(hit.shape) != (n,) or hit.dtype not in (torch.bool, torch.uint8) or not hit.is_contiguous()
                            or not (hit.device == dev or (hit.device.type == "cpu" and hit.is_pinned()))):
        raise RuntimeError("sphere_trace_packed: hit must be a contiguous bool / uint8 [N] tensor on the rays' device or in pinned "
                           "host memory")
    if queue is None:
        queue = torch.empty(1, device=dev, dtype=torch.int32)
    opts = TraceOpts(int(num_steps), 1 if compute_normals else 0, float(step_size), float(min_dis), float(far),
                     float(normal_h))
    with torch.cuda.device(dev):
        _lib.check(lib.nglod_sphere_trace_packed(ctypes.byref(view.struct), lod, _ptr(ray_o), _ptr(ray_d), n,
                                                 ctypes.byref(opts), _ptr(packed), _ptr(hit), _ptr(queue), _ptr(stats),
                                                 _stream()),
                   "nglod_sphere_trace_packed")
    return packed


def sphere_trace_camera(view, lod, origin, cam_view, right, up, tan_half_fov, ortho, window_x, window_y, workspace, packed,
                        hit=None, hit_copy=None, num_steps=256, step_size=1.0, min_dis=0.0003, far=10.0,
                        normal_h=1.0 / (64.0 * 3.0), compute_normals=True, stats=None, queue=None):
    """One camera frame, host to host, in ONE library call (nglod_sphere_trace_camera): window (pinned host or device)
    -> rays -> packed records -> optional copy of the hit bytes, all queued on the current stream.  The caller
    synchronises.  workspace: fp32 device tensor of at least 6 W H + W + H elements."""
    lib = _lib.load()
    w, h = window_x.shape[0], window_y.shape[0]
    n, dev = w * h, workspace.device
    width = 8 if hit is None else 4
    ok_mem = lambda t: t.is_contiguous() and (t.device == dev or (t.device.type == "cpu" and t.is_pinned()))
    if workspace.dtype != torch.float32 or workspace.numel() < 6 * n + w + h or not workspace.is_contiguous() or not workspace.is_cuda:
        raise RuntimeError("sphere_trace_camera: workspace must be a contiguous fp32 device tensor of >= 6 W H + W + H elements")
    if tuple(packed.shape) != (n, width) or packed.dtype != torch.float32 or not ok_mem(packed):
        raise RuntimeError(f"sphere_trace_camera: packed must be a contiguous fp32 [W*H, {width}] tensor on the device or in pinned "
                           "host memory")
    for t_ in (hit, hit_copy):
        if t_ is not None and (tuple(t_.shape) != (n,) or t_.dtype not in (torch.bool, torch.uint8) or not ok_mem(t_)):
            raise RuntimeError("sphere_trace_camera: hit / hit_copy must be contiguous bool / uint8 [W*H] tensors on the device or in "
                               "pinned host memory")
    for t_ in (window_x, window_y):
        if t_.dtype != torch.float32 or not ok_mem(t_):
            raise RuntimeError("sphere_trace_camera: the window must be contiguous fp32, on the device or in pinned host memory")
    if queue is None:
        queue = torch.empty(1, device=dev, dtype=torch.int32)
    opts = TraceOpts(int(num_steps), 1 if compute_normals else 0, float(step_size), float(min_dis), float(far),
                     float(normal_h))
    vec = [(ctypes.c_float * 3)(*[float(c) for c in v]) for v in (origin, cam_view, right, up)]
    with torch.cuda.device(dev):
        _lib.check(lib.nglod_sphere_trace_camera(ctypes.byref(view.struct), lod, vec[0], vec[1], vec[2], vec[3],
                                                 float(tan_half_fov), 1 if ortho else 0, _ptr(window_x), _ptr(window_y), w, h,
                                                 ctypes.byref(opts), _ptr(workspace), _ptr(packed), _ptr(hit), _ptr(hit_copy),
                                                 _ptr(queue), _ptr(stats), _stream()), "nglod_sphere_trace_camera")
    return packed


def unpack_trace(packed, hit=None):
    """Views (no copies) of a packed frame: {x [N,3], depth [N,1], hit [N] bool, normal [N,3]} for 32-byte records,
    {depth, hit, normal} for 16-byte records + the separate hit flags."""
    n = packed.shape[0]
    if packed.shape[1] == 4:
        return {"depth": packed[:, 0:1], "hit": hit.view(torch.bool), "normal": packed[:, 1:4]}
    hit = packed.view(torch.uint8).view(n, 32)[:, 16].view(torch.bool)
    return {"x": packed[:, 5:8], "depth": packed[:, 0:1], "hit": hit, "normal": packed[:, 1:4]}


# --------------------------------------------------------------------------- mesh2sdf / adam
def mesh2sdf_gpu(points, mesh, force_walk=False):
    """Drop-in for `mesh2sdf.mesh2sdf_gpu(points, mesh)` (mesh2sdf_kernel.cu:895-927,1008): returns [dist [N]].
    force_walk=True takes the brute-force walk whatever the batch size (the A/B reference of the large-batch path)."""
    lib = _lib.load()
    points = _f32c(points, "points")
    mesh = _f32c(mesh, "mesh")
    n = points.shape[0]
    t = mesh.shape[0]
    dist = torch.empty(n, device=points.device, dtype=torch.float32)
    with torch.cuda.device(points.device):
        _lib.check(lib.nglod_mesh2sdf_ex(_ptr(points), n, _ptr(mesh), t, _ptr(dist), _lib.M2S_FORCE_WALK if force_walk else 0,
                                         _stream()), "nglod_mesh2sdf")
    return [dist]


def release_scratch():
    """Hand the mesh2sdf scratch pool's unused memory of the current device back to the driver."""
    _lib.check(_lib.load().nglod_release_scratch(), "nglod_release_scratch")


SAMPLE_CODES = {"rand": 0, "near": 1, "trace": 2}


def mesh_area_cdf(V, F):
    """Inclusive cumulative face areas [#F] fp32 of the mesh (V [#V,3] fp32, F [#F,3] int64), for `sample_mesh`."""
    lib = _lib.load()
    V = _f32c(V, "V")
    if F.dtype != torch.int64 or not F.is_cuda:
        raise TypeError("F: expected a CUDA int64 tensor")
    F = F.contiguous()
    cdf = torch.empty(F.shape[0], device=V.device, dtype=torch.float32)
    with torch.cuda.device(V.device):
        _lib.check(lib.nglod_mesh_area_cdf(_ptr(V), _ptr(F), F.shape[0], _ptr(cdf), _stream()), "nglod_mesh_area_cdf")
    return cdf


def sample_mesh(V, F, cdf, techniques, num_samples, variance=0.01, seed=None, return_faces=False):
    """`num_samples` points per technique ('rand' | 'near' | 'trace'), concatenated in order (point_sample.py:29-57), by
    one kernel.  `seed`: Philox key; default draws one from torch's CPU generator (so `torch.manual_seed` reproduces)."""
    import ctypes
    lib = _lib.load()
    codes = [SAMPLE_CODES[t] for t in techniques]
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    dev = V.device
    V = _f32c(V, "V")
    F = F.contiguous()
    n = len(codes) * int(num_samples)
    pts = torch.empty(n, 3, device=dev, dtype=torch.float32)
    faces = torch.empty(n, device=dev, dtype=torch.int32) if return_faces else None
    arr = (ctypes.c_int * max(len(codes), 1))(*codes)
    with torch.cuda.device(dev):
        _lib.check(lib.nglod_sample_mesh(_ptr(V), _ptr(F), F.shape[0], _ptr(cdf), arr, len(codes), int(num_samples),
                                         float(variance), int(seed), _ptr(pts), _ptr(faces), _stream()), "nglod_sample_mesh")
    return (pts, faces) if return_faces else pts


def adam_step(param, grad, exp_avg, exp_avg_sq, step, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8):
    lib = _lib.load()
    n = param.numel()
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    with torch.cuda.device(param.device):
        _lib.check(lib.nglod_adam_step(_ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), n, lr, beta1, beta2,
                                       eps, bc1, bc2, _stream()), "nglod_adam_step")


# --------------------------------------------------------------------------- renderer helpers
def generate_rays(origin, view, right, up, tan_half_fov, ortho, window_x, window_y, out=None):
    """[W*H,3] ray origins / directions (x-major) from a camera basis (host 3-vectors) and device window coordinates.
    out=(ray_o, ray_d): caller-provided contiguous fp32 device buffers."""
    lib = _lib.load()
    w, h = window_x.shape[0], window_y.shape[0]
    dev = window_x.device
    if out is None:
        ray_o = torch.empty(w * h, 3, device=dev, dtype=torch.float32)
        ray_d = torch.empty(w * h, 3, device=dev, dtype=torch.float32)
    else:
        ray_o, ray_d = out
        for t_ in (ray_o, ray_d):
            if tuple(t_.shape) != (w * h, 3) or t_.dtype != torch.float32 or t_.device != dev or not t_.is_contiguous():
                raise RuntimeError("generate_rays: out buffers must be contiguous fp32 [W*H,3] tensors on the window's device")
    vec = [(ctypes.c_float * 3)(*[float(c) for c in v]) for v in (origin, view, right, up)]
    with torch.cuda.device(dev):
        _lib.check(lib.nglod_generate_rays(vec[0], vec[1], vec[2], vec[3], float(tan_half_fov), 1 if ortho else 0,
                                           _ptr(_f32c(window_x, "window_x")), _ptr(_f32c(window_y, "window_y")), w, h,
                                           _ptr(ray_o), _ptr(ray_d), _stream()), "nglod_generate_rays")
    return ray_o, ray_d


def shade_matcap(view, normal, hit, matcap, out=None):
    """In place on `normal` (misses -> 1); returns rgb with view's shape.  matcap: [U,V,C>=3] fp32 on the device.
    out: caller-provided contiguous fp32 buffer shaped like view."""
    lib = _lib.load()
    shape = view.shape
    v = _f32c(view, "view").reshape(-1, 3)
    if not (normal.is_contiguous() and normal.dtype == torch.float32):
        raise RuntimeError("normal must be a contiguous fp32 tensor (shaded in place)")
    h = hit.reshape(-1).contiguous()
    tex = _f32c(matcap, "matcap")
    if out is None:
        rgb = torch.empty_like(v)
    else:
        if out.numel() != v.numel() or out.dtype != torch.float32 or out.device != v.device or not out.is_contiguous():
            raise RuntimeError("shade_matcap: out must be a contiguous fp32 device buffer shaped like view")
        rgb = out
    with torch.cuda.device(v.device):
        _lib.check(lib.nglod_shade_matcap(_ptr(v), _ptr(normal), _ptr(h), _ptr(tex), tex.shape[0], tex.shape[1],
                                          tex.shape[2], v.shape[0], _ptr(rgb), _stream()), "nglod_shade_matcap")
    return rgb.reshape(shape)
