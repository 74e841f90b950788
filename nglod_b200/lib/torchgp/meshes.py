"""Procedural watertight test meshes (no assets ship with the reference; README.md:78-90 downloads them)."""
import math

import torch


def icosphere(subdiv=4, radius=1.0):
    """Icosahedron subdivided `subdiv` times; V [10*4^s+2, 3] float32, F [20*4^s, 3] int64, outward winding."""
    t = (1.0 + math.sqrt(5.0)) / 2.0
    verts = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
             (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    faces = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2),
             (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11),
             (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    verts = [tuple(c / math.sqrt(1 + t * t) for c in v) for v in verts]
    for _ in range(subdiv):
        cache, new_faces = {}, []

        def mid(a, b):
            key = (a, b) if a < b else (b, a)
            if key not in cache:
                m = [(verts[a][k] + verts[b][k]) / 2.0 for k in range(3)]
                n = math.sqrt(sum(c * c for c in m))
                verts.append(tuple(c / n for c in m))
                cache[key] = len(verts) - 1
            return cache[key]

        for a, b, c in faces:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            new_faces += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        faces = new_faces
    V = torch.tensor(verts, dtype=torch.float32) * radius
    F = torch.tensor(faces, dtype=torch.int64)
    return V, F


def torus(major=0.6, minor=0.25, nu=128, nv=64):
    """Torus around the y axis; nu x nv quads split into triangles; outward winding."""
    u = torch.arange(nu, dtype=torch.float64) * (2 * math.pi / nu)
    v = torch.arange(nv, dtype=torch.float64) * (2 * math.pi / nv)
    uu, vv = torch.meshgrid(u, v, indexing="ij")
    ring = major + minor * torch.cos(vv)
    V = torch.stack([ring * torch.cos(uu), minor * torch.sin(vv), ring * torch.sin(uu)], dim=-1).reshape(-1, 3)
    i = torch.arange(nu).unsqueeze(1)
    j = torch.arange(nv).unsqueeze(0)
    a = (i * nv + j).reshape(-1)
    b = (((i + 1) % nu) * nv + j).reshape(-1)
    c = (((i + 1) % nu) * nv + (j + 1) % nv).reshape(-1)
    d = (i * nv + (j + 1) % nv).reshape(-1)
    F = torch.cat([torch.stack([a, d, c], dim=1), torch.stack([a, c, b], dim=1)], dim=0)
    return V.float(), F.long()


def torus_sdf(p, major=0.6, minor=0.25):
    """Analytic SDF of the torus above (for sanity checks / synthetic targets)."""
    q = torch.sqrt(p[..., 0] ** 2 + p[..., 2] ** 2) - major
    return torch.sqrt(q * q + p[..., 1] ** 2) - minor
