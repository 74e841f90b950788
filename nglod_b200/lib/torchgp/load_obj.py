"""Minimal Wavefront OBJ reader/writer: geometry only.

The reference loads OBJs through tinyobjloader (sdf-net/lib/torchgp/load_obj.py:60-125), which is
not a dependency of the hot path; textures/materials are out of scope, so `load_materials=True`
is rejected instead of silently ignored."""
import torch


def load_obj(fname, load_materials=False):
    if load_materials:
        raise NotImplementedError("material / texture loading is outside the hot path (no tinyobjloader)")
    verts, faces = [], []
    with open(fname, "r") as fh:
        for line in fh:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "v":
                verts.append([float(tok[1]), float(tok[2]), float(tok[3])])
            elif tok[0] == "f":
                idx = []
                for item in tok[1:]:
                    i = int(item.split("/")[0])
                    idx.append(i - 1 if i > 0 else len(verts) + i)
                for k in range(1, len(idx) - 1):          # fan-triangulate polygons
                    faces.append([idx[0], idx[k], idx[k + 1]])
    V = torch.tensor(verts, dtype=torch.float32).reshape(-1, 3)
    F = torch.tensor(faces, dtype=torch.int64).reshape(-1, 3)
    return V, F


def write_obj(fname, V, F):
    with open(fname, "w") as fh:
        for v in V.tolist():
            fh.write("v {:.9g} {:.9g} {:.9g}\n".format(*v))
        for f in F.tolist():
            fh.write("f {} {} {}\n".format(f[0] + 1, f[1] + 1, f[2] + 1))
