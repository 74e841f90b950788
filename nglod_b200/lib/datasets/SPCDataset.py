"""SPCDataset -- SDF training samples for a natively sparse model (reference: sdf-net/app/spc/SPCDataset.py:51-224).

Same protocol as the reference: `init(block_idx)` draws `samples_per_voxel` uniform points inside every occupied voxel of
the finest level (sample_spc, lib/torchgp/sample_spc.py:26-43) plus as many near-surface (sigma = 2^-level) and
on-surface points again, labels them all with mesh2sdf and keeps the ones that fall into occupied voxels;
`resample(lod, idx)` selects the Morton block `idx` when the level is finer than `block_res`, shuffles and truncates to
`num_samples`.  Differences: no Kaolin (octree queries / Morton codes from lib/spc.py), no multiprocessing detour around
the SDF library (the labels come from the mesh2sdf kernel in chunks of 10^7 points), samples stay on the device unless
`to_cpu=True`, and the mesh can be handed in as `mesh=(V, F)`.
"""
import torch
from torch.utils.data import Dataset

from .. import spc as S
from ..torchgp import compute_sdf, normalize, sample_near_surface, sample_spc, sample_surface
from ..utils import setparam

_CHUNK = 10 ** 7


class SPCDataset(Dataset):
    def __init__(self, net, args=None, dataset_path=None, raw_obj_path=None, sample_mode=None, get_normals=None, seed=None,
                 num_samples=None, trim=None, samples_per_voxel=None, block_res=None, mesh=None, sdf_fn=None, to_cpu=False):
        self.dataset_path = setparam(args, dataset_path, "dataset_path")
        self.sample_mode = setparam(args, sample_mode, "sample_mode")
        self.get_normals = setparam(args, get_normals, "get_normals")
        self.num_samples = setparam(args, num_samples, "num_samples")
        self.raw_obj_path = setparam(args, raw_obj_path, "raw_obj_path")
        self.samples_per_voxel = setparam(args, samples_per_voxel, "samples_per_voxel")
        self.block_res = setparam(args, block_res, "block_res")
        self.block_size = 2 ** (self.block_res * 3)
        self.net = net
        self.to_cpu = to_cpu
        if mesh is not None:
            V, F = mesh
            dev = net.spc.octree.device
            self.V, self.F = normalize(V.float().to(dev), F.long().to(dev))
        else:
            self.V, self.F = net.V, net.F                      # the reference keeps the normalised mesh on the model
        self._sdf = sdf_fn if sdf_fn is not None else compute_sdf
        self.pts_ = self.d_ = self.pidx = self.pts = self.d = None

    # ---- Morton blocks (levels finer than block_res are trained one block of 2^(3 block_res) voxels at a time)
    def get_block(self, block_idx, mortons, subsample_res):
        """Mask of the voxels whose Morton code lies in block `block_idx` (SPCDataset.py:77-85)."""
        subsample_size = 2 ** subsample_res
        lo = self.block_size * min(block_idx, subsample_size - 1)
        return (mortons >= lo) & (mortons < lo + self.block_size)

    def get_block_idxes(self, lod=0):
        """Indices of the blocks that hold at least one voxel of LOD `lod` (SPCDataset.py:87-105)."""
        level = lod + self.net.base_lod
        mortons = S.points_to_morton(self.net.spc.level_points(level)[:, :3])
        subsample_res = 3 * max(0, level - self.block_res)
        return [i for i in range(2 ** subsample_res) if bool(self.get_block(i, mortons, subsample_res).any())]

    def _label(self, pts):
        return torch.cat([self._sdf(self.V, self.F, c) for c in torch.split(pts, _CHUNK)], dim=0)

    def init(self, block_idx=15):
        """Draw and label the sample pool (SPCDataset.py:107-169)."""
        spc = self.net.spc
        level = self.net.base_lod + self.net.num_lods - 1
        corners = spc.level_points(level)[:, :3]                 # the low corner of every occupied voxel
        subsample_res = 3 * max(0, level - self.block_res)
        if subsample_res > 0:
            corners = corners[self.get_block(block_idx, S.points_to_morton(corners), subsample_res)]
        pts = sample_spc(corners, level, self.samples_per_voxel)
        aux = []
        total = pts.shape[0]
        for size in [_CHUNK] * (total // _CHUNK) + ([total % _CHUNK] if total % _CHUNK else []):
            aux.append(sample_near_surface(self.V, self.F, size, variance=1.0 / (2 ** level)))
            aux.append(sample_surface(self.V, self.F, size)[0])
        pts = torch.cat([pts] + aux, dim=0)
        d = self._label(pts)[..., None]
        # only samples inside occupied voxels can be evaluated by the sparse model
        self.pidx = spc.query(S.quantize_points(pts, level), level)
        keep = self.pidx > -1
        self.pts_, self.d_ = pts[keep], d[keep]

    def resample(self, lod=0, idx=0):
        """Select the block, shuffle, truncate (SPCDataset.py:172-207)."""
        level = lod + self.net.base_lod
        subsample_res = 3 * max(0, level - self.block_res)
        if subsample_res > 0:
            valid_pidx = self.pidx[self.pidx > -1].clone()
            # as in the reference: the Morton codes are those of the pool's (finest-level) voxels, whatever `lod` is --
            # the trainer calls this with the finest trained LOD (main_spc.py:185)
            finest = self.net.base_lod + self.net.num_lods - 1
            vox = self.net.spc.level_points(finest)[valid_pidx.long(), :3]
            active = self.get_block(idx, S.points_to_morton(vox), subsample_res)
            self.pts, self.d = self.pts_[active], self.d_[active]
        else:
            self.pts, self.d = self.pts_, self.d_
        perm = torch.randperm(self.pts.shape[0], device=self.pts.device)
        self.pts, self.d = self.pts[perm][:self.num_samples], self.d[perm][:self.num_samples]
        if self.to_cpu:
            self.pts, self.d = self.pts.cpu(), self.d.cpu()

    def __getitem__(self, idx):
        if self.get_normals:
            return self.pts[idx], self.d[idx], self.nrm[idx]
        return self.pts[idx], self.d[idx]

    def __len__(self):
        return self.pts.shape[0]

    def num_shapes(self):
        return 1
