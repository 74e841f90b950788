"""Sparse path pinned to the reference's OWN CUDA code, compiled unmodified into oracle/_ref by
oracle/build_ref.py:

  ref_spc.spc_raytrace      sol-renderer/include/spc/spc/spc_raytrace_cuda.cpp:141-199 + spc_raytrace_cuda_kernel.cu:51-265
  ref_solr.so (ctypes)      sol-renderer/include/solr/solr/gfx/ray_aabb.cuh:104-192, sdf/sparse_grid_sample.cuh:31-109,
                            sdf/step.cuh:31-86, sdf/index_trinket.cuh:30-99, common/normalize.cuh:28-47

The host loop that strings the solr kernels together (SDF::sphereTrace / getNormal, sol-renderer/SDF.cu:218-472) is torch
C++ code inside the GL renderer and cannot be built here; `ref_sphere_trace` below restates that loop line by line and
calls the reference's compiled kernels for every device step (the decoder is torch.addmm, as in SDF.cu:412-413).
Both the CUDA product path AND the C/torch oracle restatement (oracle/) are compared with the reference here, which is
what pins the oracle the CPU tests use."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch

from oracle import nglod_oracle as O
from nglod_b200.lib import spc as S

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import build_ref  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref_spc():
    mod = build_ref.load_ref("ref_spc")
    assert mod is not None, "oracle/_ref/ref_spc/ref_spc.so is missing: run `python oracle/build_ref.py` where /root/reference exists"
    return mod


@pytest.fixture(scope="module")
def solr():
    lib = build_ref.load_solr()
    assert lib is not None, "oracle/_ref/ref_solr/ref_solr.so is missing: run `python oracle/build_ref.py`"
    assert lib.ref_solr_sizeof_trinket() == 36
    return lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _ok(rc, what):
    assert rc == 0, f"{what}: CUDA error {rc}"


def _sphere_octree(level, r=0.6):
    from test_spc import _sphere_spc
    return _sphere_spc(level, r)


def _rays(w, h, seed, extra=501):
    torch.manual_seed(seed)
    ro, rd = O.look_at([-2.8, 2.8, -2.8], [0, 0, 0], w, h, fov=30.0)
    eo = torch.rand(extra, 3) * 2 - 1                                       # origins inside the volume, random directions
    ed = torch.nn.functional.normalize(torch.randn(extra, 3), dim=1)
    return torch.cat([ro, eo]).contiguous(), torch.cat([rd, ed]).contiguous()


def _ref_raytrace(ref_spc, spc, ro, rd, target):
    """The reference's own spc_raytrace.  When targetLevel == the octree's level, its d_Decide reads `D[offset + pidx]` with
    offset = PyramidSum[Level] = the SIZE of D (spc_raytrace_cuda_kernel.cu:126: the leaf level has no child-mask bytes; the
    value is unused there, `info = notDone ? dd : 1`) -- up to 4 bytes x #leaf voxels past the end of a buffer the reference
    allocates itself, which faults or not depending on where torch's allocator happened to put it (seen as an illegal memory
    access late in the full suite, never in isolation).  The harness therefore hands the reference the same octree extended
    by ONE level in which every leaf has a single child: the read stays inside D, and nuggets at levels <= Level do not
    depend on anything below them."""
    octree, points, pyramid = spc.octree, spc.points, spc.pyramid.to(torch.int32)
    if int(target) == spc.level:
        L = spc.level
        leaf = points[int(pyramid[1, L]):int(pyramid[1, L + 1])]
        child = leaf.clone()
        child[:, :3] *= 2                                                   # child 0 of every leaf voxel
        octree = torch.cat([octree, torch.ones(leaf.shape[0], dtype=torch.uint8, device=octree.device)])
        points = torch.cat([points, child])
        ext = torch.zeros(2, L + 3, dtype=torch.int32)
        ext[0, :L + 1] = pyramid[0, :L + 1]
        ext[0, L + 1] = leaf.shape[0]
        ext[1, :L + 2] = pyramid[1, :L + 2]
        ext[1, L + 2] = pyramid[1, L + 1] + leaf.shape[0]
        pyramid = ext
    pyr = pyramid.contiguous().cpu()
    return ref_spc.spc_raytrace(octree.contiguous(), points.contiguous(), pyr, ro.contiguous(), rd.contiguous(), int(target))


# ---------------------------------------------------------------------------------------------------- a17: traversal
@pytest.mark.parametrize("level,target,res", [(5, 5, (160, 90)), (6, 4, (128, 72)), (7, 7, (320, 180)), (7, 3, (64, 36))])
def test_traversal_nuggets_equal_reference(ref_spc, level, target, res):
    """nglod_spc_raytrace_{count,fill} == the reference's level-synchronous spc_raytrace, nugget for nugget, order included;
    the C oracle likewise (this is the check that pins oracle/oracle.c's traversal)."""
    dev = "cuda"
    octree = _sphere_octree(level)
    spc = S.SPC(octree.to(dev))
    ro, rd = _rays(*res, seed=level)
    ref = _ref_raytrace(ref_spc, spc, ro.to(dev), rd.to(dev), target)
    nug, offsets = spc.raytrace(ro.to(dev), rd.to(dev), target, return_offsets=True)
    assert ref.shape[0] > 1000
    assert torch.equal(nug, ref.to(torch.int32)), (level, target)
    points, pyramid, prefix = S.octree_to_spc(octree)
    onug, ocounts = O.spc_raytrace(octree, prefix, points, pyramid, target, ro, rd)
    assert torch.equal(onug, ref.cpu().to(onug.dtype))
    assert torch.equal((offsets[1:] - offsets[:-1]).cpu(), ocounts)


def test_traversal_of_a_mesh_octree_equals_reference(ref_spc):
    """BASELINE config 4's geometry: level-7 octree of the procedural torus (mesh_to_octree), 1080p-shaped ray fan."""
    from nglod_b200.lib.torchgp import torus, normalize
    dev = "cuda"
    V, F = normalize(*[t.to(dev) for t in torus(0.6, 0.25, 128, 64)])
    torch.manual_seed(3)
    spc = S.SPC(S.mesh_to_octree(V, F, 7, num_samples=1 << 21))
    ro, rd = _rays(480, 270, seed=11, extra=2000)
    for target in (7, 5):
        ref = _ref_raytrace(ref_spc, spc, ro.to(dev), rd.to(dev), target)
        nug = spc.raytrace(ro.to(dev), rd.to(dev), target)
        assert ref.shape[0] > 50000
        assert torch.equal(nug, ref.to(torch.int32))
    # empty input / rays that miss everything
    up = torch.tensor([[0.0, 5.0, 0.0]], device=dev).repeat(64, 1)
    assert spc.raytrace(up, torch.nn.functional.normalize(torch.tensor([[0.3, 1.0, 0.2]], device=dev).repeat(64, 1), dim=1), 7).shape[0] == 0


# ---------------------------------------------------------------------------------------------------- a17: first voxel
def _ref_ray_aabb(solr, nug, points3, level, ro, rd, query, x, t, cond, pidx, init):
    """solr::ray_aabb_kernel as SDF.cu launches it (:355-372 init, :442-460 in the loop).  Mutates x, t, cond, pidx."""
    info = torch.ones(nug.shape[0], dtype=torch.int32, device=nug.device)
    if nug.shape[0] > 1:
        info[1:] = (nug[1:, 0] != nug[:-1, 0]).int()                       # d_MarkUniqueRays, sdfRenderer.cu:108-120
    info_idxes = torch.nonzero(info)[:, 0].int().contiguous()              # SDF.cu:331
    ray_inv = (1.0 / rd).contiguous()                                      # SDF.cu:322
    r = 1.0 / float(2 ** level)                                            # SDF.cu:337-338 (voxel_radius, float)
    _ok(solr.ref_solr_ray_aabb(_p(ro), _p(rd), _p(ray_inv), _p(query), _p(nug), _p(points3), _p(info), _p(info_idxes),
                               ctypes.c_float(r), 1 if init else 0, _p(x), _p(t), _p(cond), _p(pidx), nug.shape[0],
                               info_idxes.shape[0], 256 if init else 128), "ref ray_aabb_kernel")


@pytest.mark.parametrize("level,target", [(5, 5), (6, 4), (7, 7)])
def test_ray_aabb_equals_reference_kernel_bit_for_bit(solr, level, target):
    dev = "cuda"
    octree = _sphere_octree(level)
    spc = S.SPC(octree.to(dev))
    ro, rd = _rays(200, 120, seed=20 + level)
    ro, rd = ro.to(dev), rd.to(dev)
    nug, offsets = spc.raytrace(ro, rd, target, return_offsets=True)
    pts3 = spc.level_points(target)[:, :3].contiguous()
    n = ro.shape[0]
    # init: query = ray origins
    rx, rt = ro.clone(), torch.zeros(n, 1, device=dev)
    rcond, rpidx = torch.zeros(n, dtype=torch.bool, device=dev), torch.full((n,), -1, dtype=torch.int32, device=dev)
    _ref_ray_aabb(solr, nug, pts3, target, ro, rd, ro, rx, rt, rcond, rpidx, init=True)
    x, t, cond, pidx = S.ray_aabb(spc, nug, offsets, ro, rd, target)
    assert int(rcond.sum()) > 1000
    assert torch.equal(cond, rcond) and torch.equal(pidx, rpidx)
    assert torch.equal(t.view(torch.int32), rt.view(torch.int32))
    assert torch.equal(x.view(torch.int32), rx.view(torch.int32))
    # and the C oracle restatement (pins oracle/oracle.c: spc_ray_aabb)
    ox, ot, ocond, opidx = O.spc_ray_aabb(nug.cpu(), spc.level_points(target).cpu(), target, ro.cpu(), rd.cpu())
    assert torch.equal(ocond, rcond.cpu()) and torch.equal(opidx, rpidx.cpu())
    assert torch.equal(ot.view(torch.int32), rt.cpu().view(torch.int32)) and torch.equal(ox.view(torch.int32), rx.cpu().view(torch.int32))
    # re-location inside the march loop: query = a point further along the ray, only rays still alive
    g = torch.Generator(device=dev).manual_seed(5)
    adv = torch.rand(n, 1, device=dev, generator=g) * (3.0 / (1 << target))
    t2 = rt + adv * rcond.unsqueeze(1)
    q2 = torch.addcmul(ro, rd, t2).contiguous()
    rx2, rt2, rcond2, rpidx2 = q2.clone(), t2.clone(), rcond.clone(), rpidx.clone()
    _ref_ray_aabb(solr, nug, pts3, target, ro, rd, q2, rx2, rt2, rcond2, rpidx2, init=False)
    x2, tt2, cond2, pidx2 = S.ray_aabb(spc, nug, offsets, ro, rd, target, query=q2, active=rcond, t=t2)
    alive = rcond
    assert torch.equal(cond2[alive], rcond2[alive])
    ok = alive & rcond2
    assert torch.equal(pidx2[ok], rpidx2[ok])
    assert torch.equal(tt2[alive].view(torch.int32), rt2[alive].view(torch.int32))
    assert torch.equal(x2[alive].view(torch.int32), rx2[alive].view(torch.int32))


# ---------------------------------------------------------------------------------------------------- a18: sparse features
class RefTables:
    """Our sparse tables re-expressed in the reference renderer's conventions (SDF.cu:141-216): trinkets over ALL octree
    voxels (levels 0..L, global index), corner j = 4 bx + 2 by + bz (index_trinket.cuh:75-78), parent = global voxel index,
    cc = integer corner coordinates per LOD resolution, m_pyramid = inclusive cumulative voxel counts, m_res = 4 * 2^i."""

    def __init__(self, sp):
        spc, dev = sp.spc, sp.corner_feats.device
        L = spc.level
        pyr = spc.pyramid
        self.nl = sp.num_lods
        assert sp.base_lod == 2
        total = int(pyr[1, L + 1])
        tr = torch.zeros(total, 9, dtype=torch.int32, device=dev)
        tr[:, 8] = -1
        perm = [(j >> 2) | (((j >> 1) & 1) << 1) | ((j & 1) << 2) for j in range(8)]     # ours[bx + 2by + 4bz] -> ref[4bx + 2by + bz]
        nc = sp.corner_feats.shape[0]
        cc = torch.zeros(nc, 3, dtype=torch.int32, device=dev)
        for l in range(sp.num_lods):
            g0 = int(pyr[1, l + 2])
            v = slice(sp.lod_offset[l], sp.lod_offset[l + 1])
            ours = sp.trinkets[v]
            tr[g0:g0 + ours.shape[0], :8] = ours[:, perm]
            if l > 0:
                tr[g0:g0 + ours.shape[0], 8] = sp.parents[v] - sp.lod_offset[l - 1] + int(pyr[1, l + 1])
            vox = sp.voxels[v, :3].int()
            for k in range(8):
                off = torch.tensor([k & 1, (k >> 1) & 1, (k >> 2) & 1], dtype=torch.int32, device=dev)
                cc[ours[:, k].long()] = vox + off
        self.trinkets = tr.contiguous()
        self.cc = cc.contiguous()
        self.cf = sp.corner_feats.detach().float().contiguous()
        self.m_pyramid = torch.cumsum(pyr[0, :L + 1].long(), 0).to(torch.int32).to(dev).contiguous()
        self.m_res = torch.tensor([4 * 2 ** i for i in range(L - 1)], dtype=torch.int32, device=dev)
        self.w0t = [sp._decoder_params(i)[0].detach().t().contiguous() for i in range(sp.num_lods)]   # SDF.cu:81 (transpose)
        self.b0 = [sp._decoder_params(i)[1].detach() for i in range(sp.num_lods)]
        self.w1t = [sp._decoder_params(i)[2].detach().t().contiguous() for i in range(sp.num_lods)]
        self.b1 = [sp._decoder_params(i)[3].detach() for i in range(sp.num_lods)]

    def sample(self, solr, x, pidx_full, idxes, lod):
        """xs [n_active, 35] = sparse_grid_sample_kernel(x, pidx, active_idxes, ...), SDF.cu:394-410."""
        xs = torch.zeros(idxes.shape[0], 35, device=x.device)
        _ok(solr.ref_solr_sparse_grid_sample(_p(x), _p(pidx_full), _p(idxes), _p(self.trinkets), _p(self.cf), _p(self.m_pyramid),
                                             _p(self.m_res), _p(xs), idxes.shape[0], self.cf.shape[0], 32, self.nl, int(lod),
                                             _p(self.cc)), "ref sparse_grid_sample_kernel")
        return xs

    def decode(self, xs, lod):
        h = torch.relu(torch.addmm(self.b0[lod], xs, self.w0t[lod]))                     # SDF.cu:412
        return torch.addmm(self.b1[lod], h, self.w1t[lod])                               # SDF.cu:413


def _fit3_sparse(fit3):
    from test_spc import _fit3_sparse as f
    return f(fit3, "cuda")


def test_index_trinket_kernel_rebuilds_our_tables(solr, fit3):
    """The reference derives trinkets + parents from (points, cc) by brute-force search (index_trinket.cuh, driven by
    SDF::initTrinkets, SDF.cu:141-216); our sort/unique/searchsorted tables must be the same mapping."""
    net, args, spc, sp = _fit3_sparse(fit3)
    rt = RefTables(sp)
    dev = sp.corner_feats.device
    pyr = spc.pyramid
    L = spc.level
    out = torch.full_like(rt.trinkets, -7)
    pts = spc.points.contiguous()                                            # ushort4 per voxel, all levels
    counts = []
    for l in range(sp.num_lods):
        v = slice(sp.lod_offset[l], sp.lod_offset[l + 1])
        counts.append(int(sp.trinkets[v].max()) + 1 - sum(counts))
    offset_cf = 0
    for i in range(2, L + 1):                                                # SDF.cu:160-205
        parent_i = max(i - 1, 2)
        _ok(solr.ref_solr_index_trinkets(_p(pts), _p(rt.cc), _p(rt.cf), _p(out), counts[i - 2], offset_cf, int(pyr[0, i]),
                                         int(pyr[1, i]), int(pyr[0, parent_i]), int(pyr[1, parent_i]), i), "ref index_trinket_kernel")
        offset_cf += counts[i - 2]
    g0 = int(pyr[1, 2])
    assert torch.equal(out[g0:], rt.trinkets[g0:])


@pytest.mark.parametrize("math_mode,sum_lods", [("fp32", False), ("tc", False), ("tc", True), ("fp32", True)])
def test_sparse_sdf_equals_reference_sampling_kernel(solr, fit3, math_mode, sum_lods):
    """nglod_sparse_sdf_forward vs the reference's sparse_grid_sample_kernel + addmm decoder on identical points."""
    from test_spc import _points_in_voxels
    net, args, spc, sp = _fit3_sparse(fit3)
    net.math_mode = sp.math_mode = math_mode
    sp.sum_lods = sum_lods
    rt = RefTables(sp)
    osn = O.OracleSparseNet(sp.corner_feats.cpu(), sp.trinkets.cpu(), sp.parents.cpu(), sp.voxels.cpu(), sp.lod_offset, sp.base_lod,
                            [tuple(p.detach().cpu() for p in net.decoder_params(i)) for i in range(3)])
    for lod, count, spread in ((2, 50001, 1.0), (1, 1000, 1.0), (0, 33, 1.0), (2, 2000, 1.3)):
        x, pidx = _points_in_voxels(spc, lod + 2, count, 30 + lod, spread=spread)
        x, pidx = x.cuda().contiguous(), pidx.int().cuda().contiguous()
        idxes = torch.arange(count, dtype=torch.int32, device="cuda")
        xs = rt.sample(solr, x, pidx, idxes, lod)
        assert torch.equal(xs[:, :3], x)
        ref_d = rt.decode(xs, lod)
        with torch.no_grad():
            got = sp.sdf(x, lod, pidx)
            ora = osn.sdf(x.cpu(), lod, pidx.cpu().long())
            ofeat = osn.features(x.cpu(), lod, pidx.cpu().long())
        # the reference accumulates 8 * (lod + 1) products per channel in its own order: fp32 rounding differs at 1e-7
        assert (got - ref_d).abs().max() < 3e-6, (lod, float((got - ref_d).abs().max()))
        assert (ora - ref_d.cpu()).abs().max() < 3e-6                         # pins OracleSparseNet.sdf
        assert (ofeat - xs[:, 3:].cpu()).abs().max() < 2e-6                   # pins OracleSparseNet.features


# ---------------------------------------------------------------------------------------------------- a18: the stepped tracer
def ref_sphere_trace(solr, rt, nug, points3, ro, rd, lod, march_iter=50):
    """SDF::sphereTrace (sol-renderer/SDF.cu:297-472) + SDF::getNormal (:218-295): the host loop restated, every device
    step run by the reference's compiled kernels."""
    dev = ro.device
    nr = ro.shape[0]
    x = ro.clone()                                                          # :316
    t = torch.zeros(nr, 1, device=dev)
    d = torch.zeros(nr, 1, device=dev)
    dprev = torch.zeros(nr, 1, device=dev)
    pidx = torch.zeros(nr, 1, dtype=torch.int32, device=dev) - 1            # :327
    cond = torch.zeros(nr, dtype=torch.bool, device=dev)                    # :343
    hit = torch.zeros(nr, dtype=torch.bool, device=dev)
    level = lod + 2
    _ref_ray_aabb(solr, nug, points3, level, ro, rd, ro, x, t, cond, pidx, init=True)          # :355-372
    for _ in range(march_iter):                                             # :385
        active = torch.nonzero(cond)[:, 0].int().contiguous()               # :388
        if active.shape[0] == 0:
            break
        xs = rt.sample(solr, x, pidx, active, lod)                          # :404-419
        _d = rt.decode(xs, lod).contiguous()                                # :424-425
        _ok(solr.ref_solr_step(_p(ro), _p(rd), _p(active), _p(_d), _p(x), _p(t), _p(d), _p(dprev), _p(cond), _p(hit),
                               active.shape[0]), "ref step_kernel")        # :430-441
        _ref_ray_aabb(solr, nug, points3, level, ro, rd, x, x, t, cond, pidx, init=False)      # :447-464
    # getNormal (:218-295): eps 0.001 central differences inside the SAME voxel, normalised without an epsilon
    normal = torch.ones(nr, 3, device=dev)
    act = torch.nonzero(hit)[:, 0].int().contiguous()
    if act.shape[0]:
        for i in range(3):
            e = torch.zeros(3, device=dev)
            e[i] = 0.001
            x_f, x_b = (x + e).contiguous(), (x - e).contiguous()
            df = rt.decode(rt.sample(solr, x_f, pidx, act, lod), lod)
            db = rt.decode(rt.sample(solr, x_b, pidx, act, lod), lod)
            normal[hit, i] = (df - db)[:, 0]
        _ok(solr.ref_solr_normalize(_p(act), _p(normal), act.shape[0]), "ref normalize_kernel")
    return dict(x=x, depth=t, hit=hit, normal=normal, pidx=pidx[:, 0])


@pytest.mark.parametrize("math_mode,lod", [("fp32", 2), ("tc", 2), ("tc", 1)])
def test_spc_sphere_trace_equals_reference_renderer_loop(solr, ref_spc, fit3, math_mode, lod):
    """nglod_spc_sphere_trace (ONE persistent kernel) vs the reference renderer's per-step kernels on the same nuggets.
    Bit-exact where the arithmetic is integer / comparison (hit mask up to rays whose |d| straddles a threshold by
    rounding, pidx on hits); depth within 1e-4 of the scene scale (2): 2e-4; normals within 1e-3."""
    net, args, spc, sp = _fit3_sparse(fit3)
    net.math_mode = sp.math_mode = math_mode
    rt = RefTables(sp)
    torch.manual_seed(8)
    ro, rd = O.look_at([-2.8, 2.8, -2.8], [0, 0, 0], 240, 135, fov=30.0)
    ro, rd = ro.cuda().contiguous(), rd.cuda().contiguous()
    level = lod + 2
    nug = _ref_raytrace(ref_spc, spc, ro, rd, level).to(torch.int32).contiguous()      # the reference's own nuggets
    pts3 = spc.level_points(level)[:, :3].contiguous()
    with torch.no_grad():
        ref = ref_sphere_trace(solr, rt, nug, pts3, ro, rd, lod)
        x, depth, hit, normal, pidx = sp.trace(ro, rd, lod)
    h = ref["hit"]
    assert int(h.sum()) > 1500
    mism = int((hit != h).sum())
    both = h & hit
    dd = (depth - ref["depth"]).abs()[:, 0]
    nn = (normal - ref["normal"]).abs().max(dim=1)[0]
    print(f"spc trace vs reference kernels [{math_mode}, lod {lod}]: {int(h.sum())} hits / {h.numel()} rays, {mism} hit "
          f"mismatches, depth max {float(dd[both].max()):.2e}, normal max {float(nn[both].max()):.2e} "
          f"(>1e-3: {int((nn[both] > 1e-3).sum())}), pidx mismatches on hits {int((pidx[both] != ref['pidx'][both]).sum())}")
    assert mism <= 2
    assert float((dd[both] > 2e-4).float().mean()) < 2e-3
    assert float((nn[both] > 1e-3).float().mean()) < 1e-2
    assert float((pidx[both] != ref["pidx"][both]).float().mean()) < 2e-3
    # rays without any nugget are untouched by the reference (x = origin, t = 0, no hit) and by us
    no_run = torch.ones(ro.shape[0], dtype=torch.bool, device="cuda")
    no_run[nug[:, 0].long().unique()] = False
    assert not hit[no_run].any() and (depth[no_run] == 0).all()
    # the torch/C oracle restatement agrees with the reference loop too (pins oracle.spc_sphere_trace)
    osn = O.OracleSparseNet(sp.corner_feats.cpu(), sp.trinkets.cpu(), sp.parents.cpu(), sp.voxels.cpu(), sp.lod_offset, sp.base_lod,
                            [tuple(p.detach().cpu() for p in net.decoder_params(i)) for i in range(3)])
    with torch.no_grad():
        ora = O.spc_sphere_trace(osn, lod, nug.cpu(), spc.level_points(level).cpu(), ro.cpu(), rd.cpu())
    omism = int((ora["hit"] != h.cpu()).sum())
    ob = ora["hit"] & h.cpu()
    assert omism <= 2
    assert float(((ora["depth"] - ref["depth"].cpu()).abs()[:, 0][ob] > 2e-4).float().mean()) < 2e-3
