"""Build libnglod_b200.so (the C-ABI shared library) in-tree with nvcc for sm_100a.

No torch headers are involved: the library is plain CUDA C++ behind `extern "C"`
(include/nglod_b200.h), so a full rebuild is a few seconds and the resulting
.so travels to the GPU box with the repo snapshot.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OUT = os.path.join(PKG, "libnglod_b200.so")
OBJ = os.path.join(CSRC, "_build")

SOURCES = ["aabb.cu", "sdf_forward.cu", "sdf_tc.cu", "sdf_backward.cu", "sdf_backward_tc.cu", "tracer.cu", "mesh2sdf.cu", "sample_mesh.cu", "spc.cu", "spc_trace.cu", "render.cu", "probe.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-diag-suppress", "550",
]


def _flags():
    """NVCC_FLAGS plus experiment knobs from $NGLOD_EXTRA_NVCC_FLAGS (e.g. "-DNGLOD_FWD_GROUPS=2")."""
    return NVCC_FLAGS + os.environ.get("NGLOD_EXTRA_NVCC_FLAGS", "").split()


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(PKG, "..", "include")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, name), "rb") as f:
                    h.update(name.encode())
                    h.update(f.read())
    h.update(" ".join(_flags()).encode())
    return h.hexdigest()


def build_library(force=False, verbose=False):
    """Compile every .cu and link libnglod_b200.so. Returns the path."""
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "digest.txt")
    dig = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == dig:
        return OUT
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [nvcc, *_flags(), "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    r = subprocess.run([nvcc, "-shared", "-o", OUT, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
                        "-cudart", "shared"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(dig)
    return OUT


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
