"""TEST / BASELINE INFRASTRUCTURE -- import the UNMODIFIED reference Python (sdf-net/lib) and run it.

The sources are read from /root/reference/sdf-net when it exists (the authoring container) or from the copy
oracle/build_ref.py stages under the git-ignored oracle/_ref/sdf-net (the GPU box).  Third-party packages the hot path
never touches and this image does not ship (polyscope, tinyobjloader, pyexr, moviepy, matplotlib, cv2) are satisfied by
empty stub modules.  The reference's two CUDA extension modules are bound by name, in one of two ways:

  extensions="ours"       sol_nglod / mesh2sdf -> nglod_b200/shims (INTEGRATION.md route A: the reference's classes,
                          unmodified, calling the sm_100a kernels through the C ABI)
  extensions="reference"  sol_nglod / mesh2sdf -> the reference's own kernels compiled into oracle/_ref by build_ref.py
                          (the whole reference on this GPU: the baseline bench.py reports next to ours)
  extensions="cpu"        sol_nglod.aabb -> oracle/oracle.c, PerfTimer replaced (its constructor needs a CUDA driver):
                          the reference's PyTorch path on the host cores
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
STUBS = ["polyscope", "tinyobjloader", "pyexr", "moviepy", "moviepy.editor", "matplotlib", "matplotlib.pyplot", "cv2"]


def source_dir():
    for cand in ("/root/reference/sdf-net", os.path.join(HERE, "_ref", "sdf-net")):
        if os.path.isdir(os.path.join(cand, "lib")):
            return cand
    return None


class _NoTimer:
    def __init__(self, activate=False):
        pass

    def check(self, name=None):
        pass

    def reset(self):
        pass


def import_reference(extensions="ours"):
    """Returns a namespace with the reference's parse_options, OctreeSDF, SphereTracer, Renderer, MeshDataset, look_at,
    gradient.  One binding per process: the reference caches `from sol_nglod import aabb` at import time."""
    src = source_dir()
    if src is None:
        raise RuntimeError("reference Python not found: neither /root/reference/sdf-net nor oracle/_ref/sdf-net exists "
                           "(run `python oracle/build_ref.py` where /root/reference is mounted)")
    bound = getattr(import_reference, "_bound", None)
    if bound is not None and bound != extensions:
        raise RuntimeError(f"the reference is already imported with extensions={bound!r} in this process")
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    for m in STUBS:
        sys.modules.setdefault(m, types.ModuleType(m))
    if extensions == "ours":
        from nglod_b200.shims import sol_nglod as sol, mesh2sdf as m2s
        sys.modules["sol_nglod"], sys.modules["mesh2sdf"] = sol, m2s
    elif extensions == "reference":
        sys.path.insert(0, HERE)
        import build_ref
        sol, m2s = build_ref.load_ref("ref_sol_nglod"), build_ref.load_ref("ref_mesh2sdf")
        if sol is None or m2s is None:
            raise RuntimeError("oracle/_ref/ref_sol_nglod / ref_mesh2sdf are not built (python oracle/build_ref.py)")
        sys.modules["sol_nglod"], sys.modules["mesh2sdf"] = sol, m2s
    elif extensions == "cpu":
        from oracle import nglod_oracle as O
        sol = types.ModuleType("sol_nglod")
        sol.aabb = O.aabb
        sys.modules["sol_nglod"] = sol
        sys.modules.setdefault("mesh2sdf", types.ModuleType("mesh2sdf"))
    else:
        raise ValueError(extensions)
    # the reference's package is called `lib`, like ours is not: no clash, but keep it off the front of sys.path afterwards
    sys.path.insert(0, src)
    try:
        import lib.utils as rutils
        if extensions == "cpu":
            rutils.PerfTimer = _NoTimer
        ns = types.SimpleNamespace()
        ns.parse_options = importlib.import_module("lib.options").parse_options
        ns.OctreeSDF = importlib.import_module("lib.models").OctreeSDF
        ns.SphereTracer = importlib.import_module("lib.tracer").SphereTracer
        ns.Renderer = importlib.import_module("lib.renderer").Renderer
        ns.look_at = importlib.import_module("lib.geoutils").look_at
        ns.gradient = importlib.import_module("lib.diffutils").gradient
        ns.MeshDataset = importlib.import_module("lib.datasets.MeshDataset").MeshDataset
        ns.torchgp = importlib.import_module("lib.torchgp")
        ns.source = src
    finally:
        sys.path.remove(src)
    import_reference._bound = extensions
    return ns
