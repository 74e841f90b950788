"""Point samplers over a triangle mesh.

Behavioural spec (sdf-net/lib/torchgp/): area-weighted face choice
(area_weighted_distribution.py:26-45, random_face.py:27-47), barycentric
sampling with u = sqrt(r1), v = r2 (sample_surface.py:47-50), near-surface =
surface + N(0, variance) (sample_near_surface.py:43-44), uniform = U[-1,1]^3
(sample_uniform.py:31), mesh normalisation into the unit sphere (normalize.py:24-38).
All tensors stay on V's device (the reference samples on the host and copies
500k points to the GPU for every resample, MeshDataset.py:85).

On a CUDA mesh `point_sample`, `sample_surface` and `sample_near_surface` run as ONE kernel
(`nglod_sample_mesh`, csrc/sample_mesh.cu: Philox streams keyed by a seed drawn from torch's CPU
generator, binary search in the cumulative-area table); host tensors keep the reference's torch
recipe below, which is also what the kernel's distribution tests compare against.
"""
import torch

from ... import ops


def _cdf(V, F, distrib):
    """Cumulative-area table of the kernel path; `distrib` may carry a cached one (see area_weighted_distribution)."""
    cdf = getattr(distrib, "_nglod_cdf", None)
    return cdf if cdf is not None else ops.mesh_area_cdf(V, F.long())


def per_face_normals(V, F):
    tri = V[F]
    return torch.linalg.cross(tri[:, 0] - tri[:, 1], tri[:, 1] - tri[:, 2])


def area_weighted_distribution(V, F, normals=None):
    if normals is None:
        normals = per_face_normals(V, F)
    areas = torch.norm(normals, p=2, dim=1) * 0.5
    areas = areas / (torch.sum(areas) + 1e-10)
    distrib = torch.distributions.Categorical(areas.view(-1))
    if V.is_cuda:
        distrib._nglod_cdf = ops.mesh_area_cdf(V, F.long())
    return distrib


def random_face(V, F, num_samples, distrib=None):
    if distrib is None:
        distrib = area_weighted_distribution(V, F)
    normals = per_face_normals(V, F)
    idx = distrib.sample([num_samples])
    return F[idx], normals[idx]


def sample_surface(V, F, num_samples, distrib=None):
    if V.is_cuda:
        pts, faces = ops.sample_mesh(V, F.long(), _cdf(V, F, distrib), ["trace"], num_samples, return_faces=True)
        return pts, per_face_normals(V, F)[faces.long()]
    if distrib is None:
        distrib = area_weighted_distribution(V, F)
    fidx, normals = random_face(V, F, num_samples, distrib)
    f = V[fidx]
    u = torch.sqrt(torch.rand(num_samples, device=V.device)).unsqueeze(-1)
    v = torch.rand(num_samples, device=V.device).unsqueeze(-1)
    samples = (1 - u) * f[:, 0, :] + (u * (1 - v)) * f[:, 1, :] + u * v * f[:, 2, :]
    return samples, normals


def sample_near_surface(V, F, num_samples, variance=0.01, distrib=None):
    if V.is_cuda:
        return ops.sample_mesh(V, F.long(), _cdf(V, F, distrib), ["near"], num_samples, variance=variance)
    if distrib is None:
        distrib = area_weighted_distribution(V, F)
    samples = sample_surface(V, F, num_samples, distrib)[0]
    return samples + torch.randn_like(samples) * variance


def sample_uniform(num_samples, device="cpu"):
    return torch.rand(num_samples, 3, device=device) * 2.0 - 1.0


def sample_spc(corners, level, num_samples):
    """`num_samples` uniform points inside each voxel whose low corner is in `corners` [M,3] (integer coordinates at
    `level`), in [-1,1]^3 (sample_spc.py:26-43)."""
    res = 2.0 ** level
    samples = torch.rand(corners.shape[0], num_samples, 3, device=corners.device)
    samples = (corners[..., None, :3].float() + samples).reshape(-1, 3) / res
    return samples * 2.0 - 1.0


def point_sample(V, F, techniques, num_samples):
    """`num_samples` points per technique ('trace' = on-surface, 'near', 'rand'), concatenated in order."""
    if V.is_cuda:
        known = [t for t in techniques if t in ops.SAMPLE_CODES]          # the reference skips unknown names silently
        surface = any(t != "rand" for t in known)
        return ops.sample_mesh(V, F.long(), ops.mesh_area_cdf(V, F.long()) if surface else None, known, num_samples)
    distrib = None
    if "trace" in techniques or "near" in techniques:
        distrib = area_weighted_distribution(V, F)
    out = []
    for tech in techniques:
        if tech == "trace":
            out.append(sample_surface(V, F, num_samples, distrib=distrib)[0])
        elif tech == "near":
            out.append(sample_near_surface(V, F, num_samples, distrib=distrib))
        elif tech == "rand":
            out.append(sample_uniform(num_samples, device=V.device))
    return torch.cat(out, dim=0)


def normalize(V, F):
    """Centre the bounding box at the origin and scale the farthest vertex onto the unit sphere."""
    vmax, _ = torch.max(V, dim=0)
    vmin, _ = torch.min(V, dim=0)
    V = V - (vmax + vmin) / 2.0
    scale = 1.0 / torch.sqrt(torch.max(torch.sum(V ** 2, dim=-1)))
    return V * scale, F
