"""ctypes binding of libnglod_b200.so -- the C ABI declared in include/nglod_b200.h.

No torch types cross this boundary: callers pass `tensor.data_ptr()` integers,
element counts and the raw `cudaStream_t` of torch's current stream.  There is
NO fallback: if the library is missing or a call returns non-zero, this module
raises.  (The reference's own extensions never check errors -- sol_nglod_kernel.cu,
mesh2sdf_kernel.cu -- here every return code is turned into a RuntimeError.)
"""
import ctypes
import os

MAX_LODS = 8
EXPECTED_ABI = 12        # include/nglod_b200.h NGLOD_ABI_VERSION the ctypes structs below were written against
LOSS_PER_LOD = 0x80000000
MATH_TC3XTF32 = 0
MATH_FP32 = 1
M2S_FORCE_WALK = 1
EINVAL = 10001
EUNSUPPORTED = 10002

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libnglod_b200.so")

c_void_p = ctypes.c_void_p
c_int32 = ctypes.c_int32
c_int64 = ctypes.c_int64


class NetStruct(ctypes.Structure):
    """nglod_net_t"""
    _fields_ = [
        ("num_lods", c_int32),
        ("feature_dim", c_int32),
        ("hidden_dim", c_int32),
        ("pos_invariant", c_int32),
        ("math_mode", c_int32),
        ("reserved_", c_int32),
        ("grid_res", c_int32 * MAX_LODS),
        ("grids", c_void_p * MAX_LODS),
        ("w0", c_void_p * MAX_LODS),
        ("b0", c_void_p * MAX_LODS),
        ("w1", c_void_p * MAX_LODS),
        ("b1", c_void_p * MAX_LODS),
        ("summed", c_void_p * MAX_LODS),
        ("summed_fp16", c_void_p * MAX_LODS),
    ]


class SparseNetStruct(ctypes.Structure):
    """nglod_sparse_net_t"""
    _fields_ = [
        ("num_lods", c_int32),
        ("base_lod", c_int32),
        ("feature_dim", c_int32),
        ("hidden_dim", c_int32),
        ("math_mode", c_int32),
        ("pos_invariant", c_int32),
        ("lod_voxel_offset", c_int32 * (MAX_LODS + 2)),
        ("corner_feats", c_void_p),
        ("trinkets", c_void_p),
        ("parents", c_void_p),
        ("voxels", c_void_p),
        ("w0", c_void_p * MAX_LODS),
        ("b0", c_void_p * MAX_LODS),
        ("w1", c_void_p * MAX_LODS),
        ("b1", c_void_p * MAX_LODS),
        ("corner_feats_summed", c_void_p),
    ]


class NetGradStruct(ctypes.Structure):
    """nglod_net_grad_t"""
    _fields_ = [
        ("grids", c_void_p * MAX_LODS),
        ("w0", c_void_p * MAX_LODS),
        ("b0", c_void_p * MAX_LODS),
        ("w1", c_void_p * MAX_LODS),
        ("b1", c_void_p * MAX_LODS),
        ("summed", c_void_p * MAX_LODS),
        ("scatter_scratch", c_void_p),
        ("scatter_scratch_floats", c_int64),
    ]


class TraceOpts(ctypes.Structure):
    """nglod_trace_opts_t"""
    _fields_ = [
        ("num_steps", c_int32),
        ("compute_normals", c_int32),
        ("step_size", ctypes.c_double),
        ("min_dis", ctypes.c_double),
        ("far", ctypes.c_double),
        ("normal_h", ctypes.c_double),
        ("max_ctas", c_int32),
        ("reserved_", c_int32),
    ]


# name -> (restype, argtypes); must list EVERY function include/nglod_b200.h declares
SIGNATURES = {
    "nglod_abi_version": (ctypes.c_int, []),
    "nglod_build_info": (ctypes.c_char_p, []),
    "nglod_debug_tc_gemm": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "nglod_probe_gather": (ctypes.c_int, [c_void_p, c_int32, c_int64, c_int32, c_int32, c_int32, c_int32, ctypes.c_uint32,
                                          c_void_p, c_void_p]),
    "nglod_probe_scatter": (ctypes.c_int, [c_void_p, c_int32, c_int64, c_int32, c_int32, ctypes.c_uint32, c_void_p]),
    "nglod_aabb": (ctypes.c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "nglod_sdf_forward": (ctypes.c_int, [ctypes.POINTER(NetStruct), c_int32, c_void_p, c_int64, c_void_p, c_void_p]),
    "nglod_sdf_forward_all": (ctypes.c_int, [ctypes.POINTER(NetStruct), c_void_p, c_int64, c_void_p, c_void_p]),
    "nglod_sdf_features": (ctypes.c_int, [ctypes.POINTER(NetStruct), c_int32, c_void_p, c_int64, c_void_p, c_void_p]),
    "nglod_build_summed_grid": (ctypes.c_int, [ctypes.POINTER(NetStruct), c_int32, c_void_p, c_void_p]),
    "nglod_pack_grid_fp16": (ctypes.c_int, [c_void_p, c_int32, c_void_p, c_void_p]),
    "nglod_sdf_backward": (ctypes.c_int, [ctypes.POINTER(NetStruct), c_int32, c_void_p, c_int64, c_void_p,
                                          ctypes.POINTER(NetGradStruct), c_void_p, c_void_p]),
    "nglod_sdf_train_step": (ctypes.c_int, [ctypes.POINTER(NetStruct), ctypes.c_uint32, c_void_p, c_void_p, c_int64,
                                            ctypes.c_float, ctypes.POINTER(NetGradStruct), c_void_p, c_void_p]),
    "nglod_sdf_finitediff": (ctypes.c_int, [ctypes.POINTER(NetStruct), c_int32, c_void_p, c_int64, ctypes.c_float,
                                            c_void_p, c_void_p]),
    "nglod_sphere_trace": (ctypes.c_int, [ctypes.POINTER(NetStruct), c_int32, c_void_p, c_void_p, c_int64,
                                          ctypes.POINTER(TraceOpts), c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_void_p]),
    "nglod_sphere_trace_packed": (ctypes.c_int, [ctypes.POINTER(NetStruct), c_int32, c_void_p, c_void_p, c_int64,
                                                 ctypes.POINTER(TraceOpts), c_void_p, c_void_p, c_void_p, c_void_p,
                                                 c_void_p]),
    "nglod_camera_basis": (ctypes.c_int, [c_void_p, c_void_p, c_void_p]),
    "nglod_sphere_trace_camera": (ctypes.c_int, [ctypes.POINTER(NetStruct), c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                                                 ctypes.c_float, c_int32, c_void_p, c_void_p, c_int32, c_int32,
                                                 ctypes.POINTER(TraceOpts), c_void_p, c_void_p, c_void_p, c_void_p,
                                                 c_void_p, c_void_p, c_void_p]),
    "nglod_spc_raytrace_count": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, ctypes.POINTER(c_int32), c_int32, c_int32,
                                                c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "nglod_spc_raytrace_fill": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, ctypes.POINTER(c_int32), c_int32, c_int32,
                                               c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "nglod_spc_raytrace_runs": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, ctypes.POINTER(c_int32), c_int32, c_int32,
                                               c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                               c_void_p]),
    "nglod_spc_mark_first_hit": (ctypes.c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "nglod_spc_ray_aabb": (ctypes.c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int32, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "nglod_sparse_sdf_forward": (ctypes.c_int, [ctypes.POINTER(SparseNetStruct), c_int32, c_void_p, c_void_p, c_int64,
                                                c_void_p, c_void_p]),
    "nglod_sparse_sdf_backward": (ctypes.c_int, [ctypes.POINTER(SparseNetStruct), c_int32, c_void_p, c_void_p, c_int64,
                                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "nglod_sparse_sdf_train_step": (ctypes.c_int, [ctypes.POINTER(SparseNetStruct), c_int32, c_void_p, c_void_p, c_void_p,
                                                    c_int64, ctypes.c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                                                    c_void_p, c_void_p, c_void_p]),
    "nglod_spc_sphere_trace": (ctypes.c_int, [ctypes.POINTER(SparseNetStruct), c_int32, c_void_p, c_void_p, c_void_p,
                                              c_void_p, c_int64, ctypes.POINTER(TraceOpts), c_void_p, c_void_p, c_void_p,
                                              c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "nglod_spc_sphere_trace_runs": (ctypes.c_int, [ctypes.POINTER(SparseNetStruct), c_int32, c_void_p, c_void_p, c_void_p,
                                                   c_void_p, c_void_p, c_int64, ctypes.POINTER(TraceOpts), c_void_p, c_void_p,
                                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "nglod_generate_rays": (ctypes.c_int, [ctypes.POINTER(ctypes.c_float)] * 4 + [ctypes.c_float, c_int32, c_void_p, c_void_p,
                                           c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "nglod_shade_matcap": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int64,
                                          c_void_p, c_void_p]),
    "nglod_mesh2sdf": (ctypes.c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p]),
    "nglod_mesh2sdf_ex": (ctypes.c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, ctypes.c_uint32, c_void_p]),
    "nglod_release_scratch": (ctypes.c_int, []),
    "nglod_mesh_area_cdf": (ctypes.c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "nglod_sample_mesh": (ctypes.c_int, [c_void_p, c_void_p, c_int64, c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.c_int,
                                         c_int64, ctypes.c_float, ctypes.c_uint64, c_void_p, c_void_p, c_void_p]),
    "nglod_adam_step": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, ctypes.c_float,
                                       ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                       ctypes.c_float, c_void_p]),
}

_lib = None


def load():
    """Load the shared library (once) and type every entry point. Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m nglod_b200.build` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for the NGLOD hot path.")
    import torch  # noqa: F401  -- makes sure torch's libcudart.so.12 is the one we bind to
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the header and the .so disagree
        fn.restype = res
        fn.argtypes = args
    if lib.nglod_abi_version() != EXPECTED_ABI:
        raise RuntimeError(f"{LIB_PATH} has ABI version {lib.nglod_abi_version()}, this package binds version "
                           f"{EXPECTED_ABI}: stale build -- run `python -m nglod_b200.build`")
    _lib = lib
    return lib


def check(code, what):
    if code == 0:
        return
    if code == EINVAL:
        raise RuntimeError(f"{what}: invalid argument (NGLOD_EINVAL)")
    if code == EUNSUPPORTED:
        raise RuntimeError(f"{what}: unsupported model shape -- the sm_100a kernels are built for "
                           "feature_dim=32, hidden_dim=128 (NGLOD_EUNSUPPORTED)")
    raise RuntimeError(f"{what}: CUDA error {code}")
