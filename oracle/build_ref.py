"""Recipe: compile the reference's OWN CUDA extensions, unmodified, into oracle/_ref/.

TEST INFRASTRUCTURE ONLY. The sources are compiled where they lie under
/root/reference (read-only, never copied into this repo); only the built
shared objects land in oracle/_ref/ (git-ignored, but shipped to the GPU box
by gpurun).  They give the `-m gpu` parity tests a *real* reference to compare
against on the B200:

  ref_sol_nglod.aabb            <- sdf-net/lib/extensions/sol_nglod/sol_nglod_kernel.cu:161-192
  ref_mesh2sdf.mesh2sdf_gpu     <- sdf-net/lib/extensions/mesh2sdf_cuda/mesh2sdf_kernel.cu:895-1012
  ref_spc.spc_raytrace          <- sol-renderer/include/spc/spc/spc_raytrace_cuda{.cpp:141-199,_kernel.cu:51-265}
                                   (+ oracle/ref_spc_bind.cpp: the module definition only; oracle/shim/helper_math.h stands
                                   in for the un-vendored CUDA-samples header; -DCUB_NS_QUALIFIER=::kaolin::cub is what the
                                   toolkit's CUB 2.x asks for next to the file's own CUB_NS_PREFIX)
  ref_solr.so (C ABI, ctypes)   <- oracle/ref_solr_harness.cu, which #includes sol-renderer/include/solr/solr/
                                   {gfx/ray_aabb,sdf/sparse_grid_sample,sdf/step,sdf/index_trinket,common/normalize}.cuh

The reference's setup.py pins -std=c++14, which torch 2.11 headers reject, so
we drive torch.utils.cpp_extension.load() ourselves (ninja + nvcc, sm_100).
Run here (no GPU needed, nvcc cross-compiles):  python oracle/build_ref.py
"""
import os
import sys

REF = os.environ.get("NGLOD_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")

EXTS = {
    "ref_sol_nglod": ["sdf-net/lib/extensions/sol_nglod/sol_nglod_kernel.cu"],
    "ref_mesh2sdf": ["sdf-net/lib/extensions/mesh2sdf_cuda/mesh2sdf_kernel.cu"],
    "ref_spc": ["sol-renderer/include/spc/spc/spc_raytrace_cuda_kernel.cu",
                "sol-renderer/include/spc/spc/spc_raytrace_cuda.cpp", "@ref_spc_bind.cpp"],
}
EXTRA_FLAGS = {"ref_spc": ["-DCUB_NS_QUALIFIER=::kaolin::cub"]}
SOLR_SO = os.path.join(OUT, "ref_solr", "ref_solr.so")
HERE = os.path.dirname(os.path.abspath(__file__))


def build(verbose=False):
    if not os.path.isdir(REF):
        print(f"[build_ref] {REF} absent - skipping (prebuilt files in {OUT} are used if present)")
        return False
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    from torch.utils.cpp_extension import load
    ok = True
    for name, rel in EXTS.items():
        bdir = os.path.join(OUT, name)
        os.makedirs(bdir, exist_ok=True)
        so = os.path.join(bdir, name + ".so")
        if os.path.exists(so):
            print(f"[build_ref] {so} already built")
            continue
        try:
            srcs = [os.path.join(HERE, r[1:]) if r.startswith("@") else os.path.join(REF, r) for r in rel]
            load(name=name, sources=srcs, build_directory=bdir,
                 extra_include_paths=[os.path.join(HERE, "shim"), os.path.join(REF, "sol-renderer/include/spc/spc")],
                 extra_cuda_cflags=["-O3"] + EXTRA_FLAGS.get(name, []), extra_cflags=EXTRA_FLAGS.get(name, []),
                 verbose=verbose, is_python_module=False)
            print(f"[build_ref] built {so}")
        except Exception as e:  # noqa: BLE001
            ok = False
            print(f"[build_ref] FAILED {name}: {e}")
    ok = build_solr(verbose) and ok
    ok = stage_python() and ok
    return ok


PY_STAGE = os.path.join(OUT, "sdf-net")


def stage_python():
    """The reference's host side is Python and cannot travel to the GPU box any other way: stage sdf-net/lib (unmodified)
    next to the compiled extensions, under the git-ignored oracle/_ref/.  oracle/ref_python.py imports it from there, for
    INTEGRATION.md's route A test (unmodified reference Python over our shims) and bench.py's reference-on-this-B200 leg."""
    import shutil
    src = os.path.join(REF, "sdf-net", "lib")
    dst = os.path.join(PY_STAGE, "lib")
    if os.path.isdir(dst):
        print(f"[build_ref] {dst} already staged")
        return True
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "extensions"))
    print(f"[build_ref] staged {src} -> {dst}")
    return True


def build_solr(verbose=False):
    """The sol-renderer kernels behind a C ABI: plain nvcc over oracle/ref_solr_harness.cu (no torch headers)."""
    import subprocess
    if os.path.exists(SOLR_SO):
        print(f"[build_ref] {SOLR_SO} already built")
        return True
    os.makedirs(os.path.dirname(SOLR_SO), exist_ok=True)
    cmd = ["/usr/local/cuda/bin/nvcc", "-shared", "-Xcompiler", "-fPIC", "-O3", "-w", "-gencode", "arch=compute_100,code=sm_100",
           "-I", os.path.join(REF, "sol-renderer/include/solr"), os.path.join(HERE, "ref_solr_harness.cu"), "-o", SOLR_SO]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        print(f"[build_ref] FAILED ref_solr: {r.stderr[-2000:]}")
        return False
    if verbose:
        print(r.stderr)
    print(f"[build_ref] built {SOLR_SO}")
    return True


def load_solr():
    """ctypes handle of oracle/_ref/ref_solr/ref_solr.so (None if absent)."""
    import ctypes
    import torch  # noqa: F401  (libcudart)
    if not os.path.exists(SOLR_SO):
        return None
    return ctypes.CDLL(SOLR_SO)


def load_ref(name):
    """Import a prebuilt reference extension from oracle/_ref (None if absent)."""
    import importlib.util
    import torch  # noqa: F401  (the .so links against libtorch)
    so = os.path.join(OUT, name, name + ".so")
    if not os.path.exists(so):
        return None
    spec = importlib.util.spec_from_file_location(name, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    sys.exit(0 if build(verbose="-v" in sys.argv) else 1)
