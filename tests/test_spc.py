"""SPC (sparse octree): host build/decode/query pinned against the reference's numpy implementation (lib/spc3d.py,
via tests/golden/spc.npz); ray traversal: C oracle vs brute force (CPU) and CUDA vs oracle, nugget for nugget (GPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import nglod_oracle as O
from nglod_b200.lib import spc as S

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "spc.npz")


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLDEN))


def test_morton_roundtrip_and_convention():
    p = torch.tensor([[1, 0, 0], [0, 1, 0], [0, 0, 1], [3, 5, 6], [65535, 0, 65535]])
    m = S.points_to_morton(p)
    assert m[:3].tolist() == [4, 2, 1]                     # x is the most significant bit of a triple (SPC.h:65-82)
    assert torch.equal(S.morton_to_points(m).long() & 0xFFFF, p)


def test_octree_build_matches_spc3d(gold):
    level = int(gold["level"])
    leaf = torch.from_numpy(gold["leaf_points"].astype(np.int64))
    octree = S.points_to_octree(leaf[torch.randperm(leaf.shape[0])], level)       # order must not matter
    assert np.array_equal(octree.numpy(), gold["octree"])


def test_octree_decode_and_query_match_spc3d(gold):
    level = int(gold["level"])
    points, pyramid, prefix = S.octree_to_spc(torch.from_numpy(gold["octree"]))
    assert pyramid[0, :level + 1].tolist() == gold["pyramid"].tolist()
    assert pyramid[1].tolist() == [0] + np.cumsum(gold["pyramid"]).tolist()
    assert np.array_equal(S.points_to_morton(points[:, :3]).numpy().astype(np.uint64), gold["morton_all"])
    leaf = points[int(pyramid[1, level]):, :3].long()
    ref_leaf = torch.from_numpy(gold["leaf_points"].astype(np.int64))
    assert torch.equal(leaf[S.points_to_morton(leaf).argsort()], ref_leaf[S.points_to_morton(ref_leaf).argsort()])
    o = torch.from_numpy(gold["octree"]).long()
    pop = sum(((o >> i) & 1) for i in range(8))
    assert torch.equal(prefix.long(), torch.cumsum(pop, 0) - pop)

    class _Host(S.SPC):
        def __init__(self, octree):                      # no device needed for query()
            self.octree = octree
            self.points, self.pyramid, self.prefix = S.octree_to_spc(octree)
            self.level = self.pyramid.shape[1] - 2
    spc = _Host(torch.from_numpy(gold["octree"]))
    got = spc.query(torch.from_numpy(gold["query_pts"]), level)
    assert got.tolist() == gold["query_idx"].tolist()


def _sphere_spc(level, r=0.6):
    """Voxels of a sphere shell at `level` (dense test against the analytic surface, like SPC3D.construct)."""
    n = 1 << level
    ax = torch.arange(n)
    g = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), dim=-1).reshape(-1, 3)
    lo = g.float() / n * 2 - 1
    corners = torch.stack([lo + torch.tensor([i >> 2, (i >> 1) & 1, i & 1]).float() * (2.0 / n) for i in range(8)], 0)
    d = corners.norm(dim=-1) - r
    occ = (d.min(0)[0] <= 0) & (d.max(0)[0] >= 0)
    return S.points_to_octree(g[occ], level)


def _brute_force(points_lvl, level, ro, rd):
    """float64 slab test of each ray's infinite line against every voxel of the level (what d_Decide decides)."""
    n = 1 << level
    lo = (points_lvl[:, :3].double() / n * 2 - 1)
    hi = lo + 2.0 / n
    o, d = ro.double().unsqueeze(1), rd.double().unsqueeze(1)
    with np.errstate(divide="ignore", invalid="ignore"):
        t0 = (lo.unsqueeze(0) - o) / d
        t1 = (hi.unsqueeze(0) - o) / d
    tmin = torch.minimum(t0, t1).max(dim=-1)[0]
    tmax = torch.maximum(t0, t1).min(dim=-1)[0]
    return tmin, tmax


def test_oracle_traversal_vs_brute_force():
    level = 5
    octree = _sphere_spc(level)
    points, pyramid, prefix = S.octree_to_spc(octree)
    torch.manual_seed(4)
    ro, rd = O.look_at([-2.8, 2.8, -2.8], [0, 0, 0], 40, 30, fov=30.0)
    nug, counts = O.spc_raytrace(octree, prefix, points, pyramid, level, ro, rd)
    assert int(counts.sum()) == nug.shape[0] and nug.shape[0] > 1000
    assert (nug[1:, 0] >= nug[:-1, 0]).all()                                   # sorted by ray
    lp = points[int(pyramid[1, level]):]
    tmin, tmax = _brute_force(lp, level, ro, rd)
    margin = 1e-4
    surely = (tmax - tmin) > margin                                             # clearly crossed by the line
    never = (tmin - tmax) > margin                                              # clearly missed
    got = torch.zeros(ro.shape[0], lp.shape[0], dtype=torch.bool)
    got[nug[:, 0].long(), nug[:, 1].long()] = True
    assert not (got & never).any() and (got | ~surely).all()
    # front-to-back within a ray: entry distances are non-decreasing up to voxel-size ties
    te = tmin[nug[:, 0].long(), nug[:, 1].long()]
    same = nug[1:, 0] == nug[:-1, 0]
    assert ((te[1:] - te[:-1])[same] > -(2.0 / (1 << level)) * 1.8).all()
    # coarser target level = parents of the fine nuggets
    nug3, _ = O.spc_raytrace(octree, prefix, points, pyramid, 3, ro, rd)
    fine_parent = torch.unique(torch.stack([nug[:, 0].long(), S.points_to_morton(lp[nug[:, 1].long(), :3]) >> 6], 1), dim=0)
    lp3 = points[int(pyramid[1, 3]):int(pyramid[1, 4])]
    coarse = torch.unique(torch.stack([nug3[:, 0].long(), S.points_to_morton(lp3[nug3[:, 1].long(), :3])], 1), dim=0)
    cs = set(map(tuple, coarse.tolist()))
    assert all(tuple(fp) in cs for fp in fine_parent.tolist())
    # first voxel search from the ray origin
    x, t, cond, pidx = O.spc_ray_aabb(nug, lp, level, ro, rd)
    hit_rays = torch.unique(nug[:, 0].long())
    assert cond[hit_rays].float().mean() > 0.95 and not cond[counts == 0].any()
    h = cond.nonzero()[:, 0]
    assert (x[h].abs().max(dim=1)[0] <= 1.0 + 1e-5).all()
    assert torch.allclose(x[h], ro[h] + rd[h] * t[h], atol=1e-5)


@pytest.mark.gpu
def test_cuda_traversal_matches_oracle_exactly():
    dev = "cuda"
    for level, target, (w, h) in ((5, 5, (160, 90)), (6, 4, (128, 72)), (7, 7, (320, 180))):
        octree = _sphere_spc(level)
        points, pyramid, prefix = S.octree_to_spc(octree)
        torch.manual_seed(level)
        ro, rd = O.look_at([-2.8, 2.8, -2.8], [0, 0, 0], w, h, fov=30.0)
        extra_o = torch.rand(501, 3) * 2 - 1                                   # origins inside the volume, random dirs
        extra_d = torch.nn.functional.normalize(torch.randn(501, 3), dim=1)
        ro, rd = torch.cat([ro, extra_o]), torch.cat([rd, extra_d])
        ref_nug, ref_counts = O.spc_raytrace(octree, prefix, points, pyramid, target, ro, rd)
        spc = S.SPC(octree.to(dev))
        nug, offsets = spc.raytrace(ro.to(dev), rd.to(dev), target, return_offsets=True)
        assert torch.equal(nug.cpu(), ref_nug), (level, target)
        assert torch.equal((offsets[1:] - offsets[:-1]).cpu(), ref_counts)
        info = S.mark_first_hit(nug).cpu()
        assert int(info.sum()) == int((ref_counts > 0).sum())
        lp = points[int(pyramid[1, target]):int(pyramid[1, target + 1])]
        rx, rt, rcond, rpidx = O.spc_ray_aabb(ref_nug, lp, target, ro, rd)
        x, t, cond, pidx = S.ray_aabb(spc, nug, offsets, ro.to(dev), rd.to(dev), target)
        assert torch.equal(cond.cpu(), rcond) and torch.equal(pidx.cpu(), rpidx)
        assert torch.equal(t.cpu().view(torch.int32), rt.view(torch.int32))
        assert torch.equal(x.cpu().view(torch.int32), rx.view(torch.int32))
    assert spc.raytrace(ro[:0].to(dev), rd[:0].to(dev), 7).shape == (0, 2)


@pytest.mark.gpu
def test_mesh_to_octree_on_device():
    from nglod_b200.lib.torchgp import icosphere
    V, F = icosphere(4)
    octree = S.mesh_to_octree(V.cuda(), F.cuda(), 6, num_samples=1 << 20)
    spc = S.SPC(octree)
    assert spc.level == 6
    leaf = spc.level_points(6)[:, :3].float() / 64 * 2 - 1 + 1.0 / 64           # voxel centres
    assert ((leaf.norm(dim=1) - 1).abs() < 3.0 / 64 * 1.8).all()               # all voxels hug the unit sphere
    assert 8000 < leaf.shape[0] < 40000
    q = S.quantize_points(V.cuda(), 6)
    assert (spc.query(q, 6) >= 0).all()                                        # every mesh vertex lies in an occupied voxel
