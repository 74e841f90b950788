"""MeshDataset -- SDF training samples of one mesh (reference: sdf-net/lib/datasets/MeshDataset.py:35-108).

Same constructor keywords and protocol (`resample()`, `__getitem__`, `__len__`, `num_shapes`), same sampling
recipe (`sample_mode` x `num_samples` points, default 5 x 100000).  Differences: the mesh can also be handed in
directly as `mesh=(V, F)` (procedural meshes; nothing to download), and everything -- sampling, the mesh2sdf
kernel, the resulting `pts` / `d` -- stays on the device unless `to_cpu=True` asks for the reference's host copy.
"""
import torch
from torch.utils.data import Dataset

from ..torchgp import load_obj, point_sample, sample_surface, compute_sdf, normalize
from ..utils import setparam


class MeshDataset(Dataset):
    def __init__(self, args=None, dataset_path=None, raw_obj_path=None, sample_mode=None, get_normals=None,
                 seed=None, num_samples=None, trim=None, sample_tex=None, mesh=None, device="cuda", to_cpu=False):
        self.args = args
        self.dataset_path = setparam(args, dataset_path, "dataset_path")
        self.raw_obj_path = setparam(args, raw_obj_path, "raw_obj_path")
        self.sample_mode = setparam(args, sample_mode, "sample_mode")
        self.get_normals = setparam(args, get_normals, "get_normals")
        self.num_samples = setparam(args, num_samples, "num_samples")
        self.trim = setparam(args, trim, "trim")
        self.sample_tex = setparam(args, sample_tex, "sample_tex")
        self.device = torch.device(device)
        self.to_cpu = to_cpu
        if self.sample_tex:
            raise NotImplementedError("texture sampling is outside the hot path")
        if mesh is not None:
            V, F = mesh
        else:
            V, F = load_obj(self.dataset_path)
        self.V, self.F = normalize(V.float().to(self.device), F.long().to(self.device))
        self.mesh = self.V[self.F]
        self.resample()

    def resample(self):
        """Draw a fresh point set and label it with the mesh2sdf kernel."""
        self.nrm = None
        if self.get_normals:
            self.pts, self.nrm = sample_surface(self.V, self.F, self.num_samples * 5)
        else:
            self.pts = point_sample(self.V, self.F, self.sample_mode, self.num_samples)
        self.d = compute_sdf(self.V, self.F, self.pts)[..., None]
        if self.to_cpu:
            self.d, self.pts = self.d.cpu(), self.pts.cpu()
            self.nrm = None if self.nrm is None else self.nrm.cpu()

    def __getitem__(self, idx):
        if self.get_normals:
            return self.pts[idx], self.d[idx], self.nrm[idx]
        return self.pts[idx], self.d[idx]

    def __len__(self):
        return self.pts.size()[0]

    def num_shapes(self):
        return 1
