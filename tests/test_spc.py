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


# ------------------------------------------------------------------------------------------------ sparse OctreeSDF (a18)
def _fit3_sparse(fit3, device, level=4):
    """fit3 model (3 LODs, base 2 -> levels 2..4) + the octree of its torus at level 4 + sparse tables."""
    from helpers import fit3_model
    from nglod_b200.lib.torchgp import torus, normalize
    net, args = fit3_model(fit3, device)
    V, F = torus(0.6, 0.25, 64, 32)
    # shell of voxels around the analytic torus (deterministic, CPU-friendly): all level-4 voxels whose corner SDFs straddle 0
    n = 1 << level
    ax = torch.arange(n)
    g = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), dim=-1).reshape(-1, 3)
    lo = g.float() / n * 2 - 1
    from helpers import torus_sdf
    d = torch.stack([torus_sdf(lo + torch.tensor([i >> 2, (i >> 1) & 1, i & 1]).float() * (2.0 / n)) for i in range(8)], 0)
    occ = (d.min(0)[0] <= 0.02) & (d.max(0)[0] >= -0.02)
    octree = S.points_to_octree(g[occ], level).to(device)
    spc = S.SPC(octree)
    return net, args, spc, S.SparseOctreeSDF(net, spc)


def _points_in_voxels(spc, level, count, seed, spread=1.0):
    g = torch.Generator().manual_seed(seed)
    lp = spc.level_points(level)[:, :3].cpu().float()
    pidx = torch.randint(0, lp.shape[0], (count,), generator=g)
    frac = 0.5 + (torch.rand(count, 3, generator=g) - 0.5) * spread
    x = (lp[pidx] + frac) / (1 << level) * 2 - 1
    return x, pidx


def test_oracle_sparse_equals_dense_inside_voxels(fit3):
    net, args, spc, sp = _fit3_sparse(fit3, "cpu")
    osn = O.OracleSparseNet(sp.corner_feats, sp.trinkets, sp.parents, sp.voxels, sp.lod_offset, sp.base_lod,
                            [net.decoder_params(i) for i in range(3)])
    dense = O.OracleNet(net.state_dict())
    for lod in (0, 1, 2):
        x, pidx = _points_in_voxels(spc, lod + 2, 3000, lod)
        with torch.no_grad():
            assert (osn.sdf(x, lod, pidx) - dense.sdf(x, lod=lod)).abs().max() < 2e-6
    # parents really are the enclosing voxels one level up
    v = torch.arange(sp.lod_offset[2], sp.lod_offset[3])
    assert torch.equal(sp.voxels[sp.parents[v].long(), :3].long(), sp.voxels[v, :3].long() >> 1)


@pytest.mark.gpu
@pytest.mark.parametrize("math_mode,sum_lods", [("fp32", True), ("tc", True), ("tc", False)])
def test_sparse_sdf_kernel_vs_dense_and_oracle(fit3, math_mode, sum_lods):
    """sum_lods=True: the kernels read the prefix-summed corner rows and sample only the requested LOD's voxel;
    False: they walk the parent chain like the reference.  Same oracle (which walks the chain), same tolerance."""
    net, args, spc, sp = _fit3_sparse(fit3, "cuda")
    net.math_mode = sp.math_mode = math_mode
    assert sp.corner_feats_summed is not None
    sp.sum_lods = sum_lods
    osn = O.OracleSparseNet(sp.corner_feats, sp.trinkets, sp.parents, sp.voxels, sp.lod_offset, sp.base_lod,
                            [net.decoder_params(i) for i in range(3)])
    for lod, count in ((2, 50001), (1, 1000), (0, 33)):
        x, pidx = _points_in_voxels(spc, lod + 2, count, 10 + lod)
        with torch.no_grad():
            got = sp.sdf(x.cuda(), lod, pidx.cuda()).cpu()
            assert (got - net.sdf(x.cuda(), lod=lod).cpu()).abs().max() < 3e-6
            assert (got - osn.sdf(x, lod, pidx)).abs().max() < 3e-6
    # slightly outside the voxel: weights extrapolate (not clamped), like the reference kernel
    x, pidx = _points_in_voxels(spc, 4, 2000, 99, spread=1.3)
    with torch.no_grad():
        assert (sp.sdf(x.cuda(), 2, pidx.cuda()).cpu() - osn.sdf(x, 2, pidx)).abs().max() < 3e-6
    assert sp.sdf(x[:0].cuda(), 2, pidx[:0].cuda()).shape == (0, 1)


@pytest.mark.gpu
@pytest.mark.parametrize("math_mode", ["fp32", "tc"])
def test_spc_sphere_trace_vs_oracle(fit3, math_mode):
    net, args, spc, sp = _fit3_sparse(fit3, "cuda")
    net.math_mode = sp.math_mode = math_mode
    osn = O.OracleSparseNet(sp.corner_feats, sp.trinkets, sp.parents, sp.voxels, sp.lod_offset, sp.base_lod,
                            [net.decoder_params(i) for i in range(3)])
    torch.manual_seed(8)
    ro, rd = O.look_at([-2.8, 2.8, -2.8], [0, 0, 0], 160, 90, fov=30.0)
    x, depth, hit, normal, pidx = sp.trace(ro.cuda(), rd.cuda(), 2)
    nug = spc.raytrace(ro.cuda(), rd.cuda(), 4).cpu()
    with torch.no_grad():
        ref = O.spc_sphere_trace(osn, 2, nug, spc.level_points(4).cpu(), ro, rd)
    h = ref["hit"]
    mism = int((hit.cpu() != h).sum())
    both = h & hit.cpu()
    dd = (depth.cpu() - ref["depth"]).abs()[:, 0]
    nn = (normal.cpu() - ref["normal"]).abs().max(dim=1)[0]
    print(f"spc trace [{math_mode}]: {int(h.sum())} hits of {h.numel()} rays, {mism} hit mismatches, depth max "
          f"{float(dd[both].max()):.2e}, normal max {float(nn[both].max()):.2e} (>1e-3: {int((nn[both] > 1e-3).sum())})")
    assert int(h.sum()) > 800
    assert mism <= 2
    assert float((dd[both] > 2e-4).float().mean()) < 2e-3
    assert float((nn[both] > 1e-3).float().mean()) < 1e-2
    assert (normal.cpu()[~hit.cpu()] == 0).all()
    # hits lie on the fitted surface: the dense model agrees the SDF is ~0 there
    with torch.no_grad():
        assert net.sdf(x[hit], lod=2).abs().max() < 5e-3
    no_run = torch.ones(ro.shape[0], dtype=torch.bool)
    no_run[nug[:, 0].long().unique()] = False
    assert not hit.cpu()[no_run].any() and (depth.cpu()[no_run] == 0).all()


# ------------------------------------------------------------------------------------------------ sparse training
def _shell_octree(level, device):
    n = 1 << level
    ax = torch.arange(n)
    g = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), dim=-1).reshape(-1, 3)
    lo = g.float() / n * 2 - 1
    from helpers import torus_sdf
    d = torch.stack([torus_sdf(lo + torch.tensor([i >> 2, (i >> 1) & 1, i & 1]).float() * (2.0 / n)) for i in range(8)], 0)
    occ = (d.min(0)[0] <= 0.02) & (d.max(0)[0] >= -0.02)
    return S.points_to_octree(g[occ], level).to(device)


@pytest.mark.gpu
def test_one_pass_runs_equal_the_two_pass_nugget_list(fit3):
    """nglod_spc_raytrace_runs (one pass, per-ray runs in arbitrary ray order) against the two-pass list that is pinned to the
    reference's spc_raytrace: every ray's run holds the same voxels in the same order; the tracer over the runs returns
    the two-pass frame bit for bit; a buffer that is too small is reported and the frame is redone through the list."""
    net, args, spc, sp = _fit3_sparse(fit3, "cuda")
    torch.manual_seed(8)
    ro, rd = O.look_at([-2.8, 2.8, -2.8], [0, 0, 0], 160, 90, fov=30.0)
    ro, rd = ro.cuda().contiguous(), rd.cuda().contiguous()
    nug, off = spc.raytrace(ro, rd, 4, return_offsets=True)
    rn, rb, re, cur = S._raytrace_runs(spc, ro, rd, 4)
    assert int(cur[1]) == 0 and int(cur[0]) == nug.shape[0]
    assert torch.equal((re - rb), (off[1:] - off[:-1]))
    nz = (re > rb).nonzero(as_tuple=True)[0]
    assert nz.numel() > 1000
    longest = int((re - rb).max())
    assert longest > 8                                   # runs longer than the register buffer take the second walk
    for k in range(longest):                             # k-th nugget of every run that has one
        sel = nz[(re - rb)[nz] > k]
        a = rn[(rb[sel] + k).long()]
        b = nug[(off[sel] + k).long()]
        assert torch.equal(a, b)
    f2 = sp.trace(ro, rd, 2, one_pass=False)
    f1 = sp.trace(ro, rd, 2, one_pass=True)
    for a, b in zip(f1, f2):
        assert torch.equal(a, b)
    small = sp.trace(ro, rd, 2, one_pass=True, runs_capacity=1000)      # overflows -> falls back to the list
    assert int(spc._runs_ws["cursor"][1]) == 1
    for a, b in zip(small, f2):
        assert torch.equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("pos_invariant", [False, True])
def test_neural_spc_forward_backward_vs_oracle(pos_invariant):
    """NeuralSPC (features only on the corners of occupied voxels; the reference's app/spc model): sdf and its gradients
    w.r.t. the corner features and the decoder against autograd through the oracle's parent-chain restatement."""
    spc = S.SPC(_shell_octree(5, "cuda"))
    torch.manual_seed(1)
    net = S.NeuralSPC(spc, num_lods=3, base_lod=3, feature_std=0.2, pos_invariant=pos_invariant)
    assert net.corner_feats.shape[0] == sum(net.corner_counts) and net.trinkets.shape[0] == net.lod_offset[-1]
    for lod in (2, 0):
        x, pidx = _points_in_voxels(spc, lod + 3, 4001, 20 + lod)
        assert torch.equal(net.query(x.cuda(), lod).cpu(), pidx)
        osn = O.OracleSparseNet(net.corner_feats, net.trinkets, net.parents, net.voxels, net.lod_offset, net.base_lod,
                                [net._decoder_params(i) for i in range(3)], pos_invariant=pos_invariant)
        osn.cf.requires_grad_(True)
        for t in osn.dec[lod]:
            t.requires_grad_(True)
        gt = torch.rand(x.shape[0], 1, generator=torch.Generator().manual_seed(3)) - 0.5
        ref = osn.sdf(x, lod, pidx)
        loss_ref = ((ref - gt) ** 2).mean()
        loss_ref.backward()
        for p in net.parameters():
            p.grad = None
        d = net.sdf(x.cuda(), lod, pidx.cuda())
        assert (d.detach().cpu() - ref.detach()).abs().max() < 5e-6
        loss = ((d - gt.cuda()) ** 2).mean()
        loss.backward()
        assert abs(loss.item() - loss_ref.item()) < 1e-5 * max(1.0, loss_ref.item())
        g, r = net.corner_feats.grad.cpu(), osn.cf.grad
        assert (g - r).abs().max() / r.abs().max() < 3e-4
        assert (r != 0).any(dim=1).sum() > 100 and ((g != 0).any(dim=1) == (r != 0).any(dim=1)).float().mean() > 0.999
        for a, b in zip(net._decoder_params(lod), osn.dec[lod]):
            assert (a.grad.cpu() - b.grad).abs().max() / b.grad.abs().max() < 3e-4


def test_spc_dataset_protocol_on_cpu():
    """SPCDataset (SPCDataset.py:51-224) host logic, on the CPU with the oracle's mesh2sdf as the labeller: the pool holds
    samples_per_voxel points per occupied voxel plus twice as many surface points, only samples inside occupied voxels
    survive, labels are the labeller's, blocks partition the voxels by Morton range, resample shuffles and truncates."""
    import types
    from nglod_b200.lib.datasets import SPCDataset
    from nglod_b200.lib.torchgp import torus, normalize
    level = 4
    spc = S.SPC(_shell_octree(level, "cpu"))
    net = types.SimpleNamespace(spc=spc, base_lod=2, num_lods=3)              # finest level = 2 + 3 - 1 = 4
    V, F = torus(0.6, 0.25, 24, 12)
    calls = []

    def labeller(Vn, Fn, p):
        calls.append(p.shape[0])
        return O.mesh2sdf(p, Vn[Fn])

    torch.manual_seed(0)
    ds = SPCDataset(net, args=None, sample_mode=["rand"], get_normals=False, num_samples=3000, samples_per_voxel=4,
                    block_res=7, mesh=(V, F), sdf_fn=labeller)
    assert ds.get_block_idxes(lod=2) == [0]                                   # level 4 <= block_res: one block
    ds.init()
    nvox = spc.level_points(level).shape[0]
    assert sum(calls) == 3 * 4 * nvox                                         # voxel samples + near + trace
    q = spc.query(S.quantize_points(ds.pts_, level), level)
    assert (q > -1).all() and ds.pts_.shape[0] == int((ds.pidx > -1).sum())
    assert ds.pts_.shape[0] >= 4 * nvox                                       # every in-voxel sample survives
    Vn, Fn = normalize(V, F)
    assert torch.equal(ds.d_[:500, 0], O.mesh2sdf(ds.pts_[:500], Vn[Fn]))
    ds.resample(lod=2, idx=0)
    assert len(ds) == 3000 and ds.num_shapes() == 1
    p0, d0 = ds[0]
    assert p0.shape == (3,) and d0.shape == (1,)
    first = ds.pts.clone()
    ds.resample(lod=2, idx=0)
    assert not torch.equal(first, ds.pts)                                      # a new permutation
    # finer than block_res: blocks of 2^(3 block_res) Morton codes, each resample stays inside its block
    ds2 = SPCDataset(net, args=None, sample_mode=["rand"], get_normals=False, num_samples=10 ** 9, samples_per_voxel=2,
                     block_res=3, mesh=(V, F), sdf_fn=labeller)
    blocks = ds2.get_block_idxes(lod=2)
    mort = S.points_to_morton(spc.level_points(level)[:, :3])
    assert blocks == sorted(set((mort >> 9).tolist())) and len(blocks) > 1
    ds2.init(block_idx=blocks[1])
    vq = spc.level_points(level)[spc.query(S.quantize_points(ds2.pts_, level), level), :3]
    inside = (S.points_to_morton(vq) >> 9) == blocks[1]
    assert inside.sum() >= 2 * int((mort >> 9 == blocks[1]).sum())            # the block's voxel samples (+ surface samples in it)
    ds2.resample(lod=2, idx=blocks[1])
    vq = spc.level_points(level)[spc.query(S.quantize_points(ds2.pts, level), level), :3]
    assert ((S.points_to_morton(vq) >> 9) == blocks[1]).all() and len(ds2) == int(inside.sum())


@pytest.mark.gpu
@pytest.mark.parametrize("pos_invariant", [False, True])
def test_neural_spc_fused_loss_backward_equals_autograd(pos_invariant):
    """NeuralSPC.loss_backward (one fused kernel per head: forward + L2 loss + backward, nglod_sparse_sdf_train_step) against
    the same loss through autograd (separate sparse forward, torch loss, nglod_sparse_sdf_backward): losses and every
    gradient, two heads accumulated into the same corner-feature gradient."""
    spc = S.SPC(_shell_octree(5, "cuda"))
    torch.manual_seed(4)
    net = S.NeuralSPC(spc, num_lods=3, base_lod=3, feature_std=0.2, pos_invariant=pos_invariant)
    x, _ = _points_in_voxels(spc, 5, 6007, 11)            # inside level-5 voxels, hence inside their ancestors too
    x = x.cuda()
    gt = (torch.rand(x.shape[0], 1, generator=torch.Generator().manual_seed(5)) - 0.5).cuda()
    lods = [0, 2]
    for p in net.parameters():
        p.grad = None
    ref_losses = []
    total = 0
    for lod in lods:
        l = ((net.sdf(x, lod) - gt) ** 2).sum() / x.shape[0]
        ref_losses.append(l.item())
        total = total + l
    total.backward()
    ref = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
    for p in net.parameters():
        p.grad = None
    losses = net.loss_backward(x, gt, lods=lods)
    for a, b in zip(losses.tolist(), ref_losses):
        assert abs(a - b) < 1e-5 * max(1.0, abs(b))
    got = {k: p.grad for k, p in net.named_parameters() if p.grad is not None}
    for k, r in ref.items():
        assert (got[k] - r).abs().max() / r.abs().max() < 3e-4, k
    for k, g in got.items():                              # heads that were not trained: zero gradient, not garbage
        if k not in ref:
            assert not g.any(), k
    # accumulates: a second call doubles the gradient
    net.loss_backward(x, gt, lods=lods)
    assert (net.corner_feats.grad - 2 * ref["corner_feats"]).abs().max() / ref["corner_feats"].abs().max() < 6e-4
    # points outside the octree (SPC.query -> -1) are inert rows: no table is read for them, they add no loss and no
    # gradient (the reference's Python indexing would wrap to the last voxel; reading trinkets[-8..] is not an option)
    xo = torch.cat([x, torch.rand(999, 3, device="cuda") * 2 - 1])                   # 999 extra rows marked "no voxel"
    gto = torch.cat([gt, torch.zeros(999, 1, device="cuda")])
    none = torch.full((999,), -1, dtype=torch.int64, device="cuda")
    pidx_o = [torch.cat([net.query(x, lod), none]) for lod in lods]
    d_in = net.sdf(x, 2)
    d_all = net.sdf(xo, 2, pidx_o[1])
    assert torch.equal(d_all[:x.shape[0]], d_in) and (d_all[x.shape[0]:] == 0).all()
    for p in net.parameters():
        p.grad = None
    losses_o = net.loss_backward(xo, gto, lods=lods, pidx=pidx_o, global_batch=x.shape[0])
    for a, b in zip(losses_o.tolist(), ref_losses):
        assert abs(a - b) < 1e-5 * max(1.0, abs(b))
    assert (net.corner_feats.grad - ref["corner_feats"]).abs().max() / ref["corner_feats"].abs().max() < 3e-4
    for p in net.parameters():
        p.grad = None
    ((net.sdf(xo, 2, pidx_o[1]) - gto) ** 2).sum().backward()                        # autograd path, same guard
    assert torch.isfinite(net.corner_feats.grad).all()
    g_ref = net.corner_feats.grad.clone()
    for p in net.parameters():
        p.grad = None
    ((net.sdf(x, 2) - gt) ** 2).sum().backward()
    assert (net.corner_feats.grad - g_ref).abs().max() / g_ref.abs().max() < 3e-4


@pytest.mark.gpu
def test_neural_spc_trains_and_traces():
    """A few hundred Adam steps on points inside the occupied voxels of a torus shell: the loss drops by >10x, and the
    in-voxel tracer renders the fitted surface (hits lie on the analytic torus)."""
    from helpers import torus_sdf
    spc = S.SPC(_shell_octree(5, "cuda"))
    torch.manual_seed(0)
    net = S.NeuralSPC(spc, num_lods=3, base_lod=3)
    opt = torch.optim.Adam(net.parameters(), lr=3e-3)
    lp = spc.level_points(5)[:, :3].float()
    g = torch.Generator(device="cuda").manual_seed(2)
    first = last = None
    for it in range(300):
        pidx = torch.randint(0, lp.shape[0], (8192,), device="cuda", generator=g)
        x = (lp[pidx] + torch.rand(8192, 3, device="cuda", generator=g)) / 32 * 2 - 1
        gt = torus_sdf(x).unsqueeze(1)
        opt.zero_grad(set_to_none=True)
        loss = ((net.sdf(x, 2, pidx) - gt) ** 2).mean()
        loss.backward()
        opt.step()
        first = loss.item() if first is None else first
        last = loss.item()
    print(f"NeuralSPC fit: loss {first:.3e} -> {last:.3e}")
    assert last < first / 10
    torch.manual_seed(8)
    ro, rd = O.look_at([-2.8, 2.8, -2.8], [0, 0, 0], 160, 90, fov=30.0)
    with torch.no_grad():
        x, depth, hit, normal, _ = net.trace(ro.cuda(), rd.cuda(), 2)
    assert int(hit.sum()) > 500
    assert torus_sdf(x[hit]).abs().max() < 0.03
    # inference in eval mode samples one level of the prefix-summed corner rows: same function as the parent-chain walk
    net.eval()
    xs, pidx = _points_in_voxels(spc, 5, 5000, 7)
    with torch.no_grad():
        a = net.sdf(xs.cuda(), 2, pidx.cuda())
        assert net.corner_feats_summed is not None
        net.sum_lods = False
        b = net.sdf(xs.cuda(), 2, pidx.cuda())
        assert net.corner_feats_summed is None
        net.sum_lods = True
        x2, depth2, hit2, normal2, _ = net.trace(ro.cuda(), rd.cuda(), 2)
    assert (a - b).abs().max() < 2e-6
    assert int((hit2 != hit).sum()) <= 2


# ------------------------------------------------------------------------------------------------ sparse model file (f2)
def test_sparse_model_file_roundtrip_tables(fit3, tmp_path):
    """SparseOctreeSDF.save -> .load (the real-time renderer's npz, SOL_NGLOD.py:80-100 / SDF.cu:65-216): the tables the
    reader re-derives from `cc` are the writer's tables; features and decoders come back as their fp16 roundings."""
    net, args, spc, sp = _fit3_sparse(fit3, "cpu")
    path = str(tmp_path / "torus3.npz")
    sp.save(path)
    z = np.load(path)
    assert z["cc"].dtype == np.uint8 and z["cf"].dtype == np.float16 and z["w0"].dtype == np.float16
    assert z["octree"].dtype == np.uint8 and z["pyramid"].shape == (3,) and z["w0"].shape == (3, 128, 35)
    sp2 = S.SparseOctreeSDF.load(path, device="cpu")
    assert sp2.num_lods == 3 and sp2.base_lod == 2 and sp2.lod_offset == sp.lod_offset
    assert torch.equal(sp2.trinkets, sp.trinkets) and torch.equal(sp2.parents, sp.parents)
    assert torch.equal(sp2.voxels, sp.voxels) and torch.equal(sp2.spc.octree, spc.octree)
    assert torch.equal(sp2.corner_feats, sp.corner_feats.half().float())
    for i in range(3):
        for a, b in zip(sp2._decoder_params(i), net.decoder_params(i)):
            assert torch.equal(a, b.detach().half().float())
    # a file whose corner table does not cover the octree is rejected
    bad = dict(z)
    bad["cc"] = bad["cc"].copy()
    bad["cc"][5] = [255, 255, 255]
    np.savez_compressed(str(tmp_path / "bad.npz"), **bad)
    with pytest.raises(ValueError):
        S.SparseOctreeSDF.load(str(tmp_path / "bad.npz"), device="cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("math_mode", ["fp32", "tc"])
def test_sparse_model_file_save_load_trace_identity(fit5, tmp_path, math_mode):
    """fit5's weights are fp16-exact, so the file holds exactly the model that wrote it: the loaded model traces the same
    frame bit for bit (x, depth, hit, normal, last voxel) and evaluates the same sdf."""
    from helpers import fit5_model, torus_sdf
    net, args = fit5_model(fit5, "cuda")
    net.math_mode = math_mode
    level = 6
    n = 1 << level
    ax = torch.arange(n)
    g = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), dim=-1).reshape(-1, 3)
    lo = g.float() / n * 2 - 1
    d = torch.stack([torus_sdf(lo + torch.tensor([i >> 2, (i >> 1) & 1, i & 1]).float() * (2.0 / n)) for i in range(8)], 0)
    occ = (d.min(0)[0] <= 0.01) & (d.max(0)[0] >= -0.01)
    spc = S.SPC(S.points_to_octree(g[occ], level).cuda())
    sp = S.SparseOctreeSDF(net, spc)
    sp.math_mode = math_mode
    path = str(tmp_path / "torus5.npz")
    sp.save(path)
    sp2 = S.SparseOctreeSDF.load(path, device="cuda", math_mode=math_mode)
    assert torch.equal(sp2.corner_feats, sp.corner_feats) and torch.equal(sp2.trinkets, sp.trinkets)
    torch.manual_seed(9)
    ro, rd = O.look_at([-2.8, 2.8, -2.8], [0, 0, 0], 320, 180, fov=30.0)
    for lod in (4, 2):
        a = sp.trace(ro.cuda(), rd.cuda(), lod)
        b = sp2.trace(ro.cuda(), rd.cuda(), lod)
        assert int(a[2].sum()) > 3000
        for u, v in zip(a, b):
            assert torch.equal(u, v)
        x, pidx = _points_in_voxels(spc, lod + 2, 5000, 3)
        assert torch.equal(sp.sdf(x.cuda(), lod, pidx.cuda()), sp2.sdf(x.cuda(), lod, pidx.cuda()))
    # points outside every voxel (pidx = -1): inert rows, no out-of-bounds table reads
    x, pidx = _points_in_voxels(spc, 6, 1000, 5)
    pidx = pidx.clone()
    pidx[::3] = -1
    out = sp2.sdf(x.cuda(), 4, pidx.cuda())
    assert (out[::3] == 0).all() and torch.isfinite(out).all()
    ok = pidx >= 0
    assert torch.equal(out[ok], sp2.sdf(x[ok].cuda(), 4, pidx[ok].cuda()))
