"""Sparse octree (SPC) of a mesh and ray traversal over it -- Kaolin-free.

The reference's sparse path (sdf-net/app/spc) calls un-vendored Kaolin for every primitive; each has an in-tree twin
that defines the behaviour.  This module provides them over torch tensor ops (octree
build / decode: set-up work, done once per mesh) and the sm_100a kernels (ray traversal: per frame):

  points_to_morton / morton_to_points   <- ToMorton / ToPoint, sol-renderer/include/spc/spc/SPC.h:65-97 (= lib/spc3d.py:24-40)
  quantize_points                       <- kaolin.ops.spc.quantize_points (call site spc_utils.py:79)
  points_to_octree                      <- SPC3D.points_to_nodes, sdf-net/lib/spc3d.py:224-255
  octree_to_spc                         <- SPC::SetGeometry, sol-renderer/include/spc/spc/SPC.cu:148-238 (= scan_octrees + generate_points)
  mesh_to_octree                        <- spc_utils.py:74-84
  query                                 <- SPC3D.Identify, lib/spc3d.py:309-338 (= unbatched_query)
  SPC.raytrace / mark_first_hit / ray_aabb  <- spc_raytrace, d_MarkUniqueRays, ray_aabb_kernel (see csrc/spc.cu)
"""
import ctypes

import torch

from .. import _lib, ops
from ..ops import _ptr, _stream, _f32c
from .torchgp import sample_surface


def points_to_morton(points):
    """[N,3] integer voxel coordinates -> int64 Morton codes; x is the most significant bit of each triple."""
    p = points.long()
    code = torch.zeros(p.shape[0], dtype=torch.int64, device=p.device)
    for i in range(16):
        bit = 1 << i
        code |= ((p[:, 0] & bit) << (2 * i + 2)) | ((p[:, 1] & bit) << (2 * i + 1)) | ((p[:, 2] & bit) << (2 * i))
    return code


def morton_to_points(morton):
    m = morton.long()
    p = torch.zeros(m.shape[0], 3, dtype=torch.int64, device=m.device)
    for i in range(16):
        p[:, 0] |= ((m >> (3 * i + 2)) & 1) << i
        p[:, 1] |= ((m >> (3 * i + 1)) & 1) << i
        p[:, 2] |= ((m >> (3 * i)) & 1) << i
    return p.short()


def quantize_points(x, level):
    """[-1,1]^3 -> integer voxel coordinates in [0, 2^level)."""
    res = 2 ** level
    return torch.floor(torch.clamp(res * (x + 1.0) / 2.0, 0, res - 1.0)).short()


def points_to_octree(points, level):
    """Unique voxel coordinates at `level` (any order) -> uint8 child-mask bytes, breadth first (root first)."""
    m = torch.unique(points_to_morton(points))            # sorted
    levels = []
    for _ in range(level):
        parent = m >> 3
        child = (m & 7)
        uniq, inv = torch.unique(parent, return_inverse=True)
        byte = torch.zeros(uniq.shape[0], dtype=torch.int64, device=m.device)
        byte.scatter_add_(0, inv, (1 << child))           # children of one parent are distinct -> sum == OR
        levels.append(byte.to(torch.uint8))
        m = uniq
    return torch.cat(levels[::-1]) if levels else torch.zeros(0, dtype=torch.uint8, device=points.device)


def octree_to_spc(octree):
    """child-mask bytes -> (points [psize,4] int16 of all levels in Morton order, pyramid [2, L+2] int32 on the
    host (row 0: voxels per level, row 1: first point of each level), prefix int32 = exclusive popcount sum)."""
    dev = octree.device
    o = octree.long()
    pop = torch.zeros_like(o)
    for i in range(8):
        pop += (o >> i) & 1
    prefix = (torch.cumsum(pop, 0) - pop).int()
    mortons = [torch.zeros(1, dtype=torch.int64, device=dev)]
    counts = [1]
    start = 0
    while start < o.shape[0]:
        n = counts[-1]
        bytes_l = o[start:start + n]
        cand = (mortons[-1].unsqueeze(1) * 8 + torch.arange(8, device=dev).unsqueeze(0))
        keep = ((bytes_l.unsqueeze(1) >> torch.arange(8, device=dev).unsqueeze(0)) & 1).bool()
        nxt = cand[keep]                                   # node-major, child index ascending == Morton order
        mortons.append(nxt)
        counts.append(int(nxt.shape[0]))
        start += n
    level = len(counts) - 1
    pts3 = morton_to_points(torch.cat(mortons))
    points = torch.zeros(pts3.shape[0], 4, dtype=torch.int16, device=dev)
    points[:, :3] = pts3
    pyramid = torch.zeros(2, level + 2, dtype=torch.int32)
    s = 0
    for l, c in enumerate(counts):
        pyramid[0, l] = c
        pyramid[1, l] = s
        s += c
    pyramid[1, level + 1] = s
    return points, pyramid, prefix


def mesh_to_octree(V, F, level, num_samples=1 << 24):
    """Octree of the voxels a mesh surface touches (spc_utils.py:74-84): surface samples plus a half-voxel jittered
    copy, quantised, de-duplicated.  The reference draws 1e8 samples; `num_samples` trades build time for coverage."""
    if V.is_cuda:           # points only: skip the per-sample normal gather of sample_surface
        samples = ops.sample_mesh(V, F.long(), ops.mesh_area_cdf(V, F.long()), ["trace"], num_samples)
    else:
        samples = sample_surface(V, F, num_samples)[0]
    samples = torch.cat([samples, samples + (torch.rand_like(samples) * 2.0 - 1.0) * (1.0 / (2 ** (level + 1)))], dim=0)
    q = quantize_points(samples, level)
    return points_to_octree(torch.unique(q.long(), dim=0), level)


class SPC:
    """Geometry of a sparse octree + ray traversal on the device (the geometric half of sdf-net/app/spc/SPC.py)."""

    def __init__(self, octree):
        self.octree = octree.contiguous()
        self.points, self.pyramid, self.prefix = octree_to_spc(self.octree)
        self.level = self.pyramid.shape[1] - 2
        self._pyrsum = (ctypes.c_int32 * (self.level + 2))(*[int(v) for v in self.pyramid[1]])

    def level_points(self, level):
        return self.points[int(self.pyramid[1, level]):int(self.pyramid[1, level + 1])]

    def query(self, qpts, level):
        """Index (within `level`) of the voxel holding each integer point, -1 if unoccupied (spc3d.py:309-338)."""
        q = qpts.long()
        n = q.shape[0]
        o = self.octree.long()
        psum = self.prefix.long()
        ordn = torch.zeros(n, dtype=torch.int64, device=q.device)
        alive = ((q >= 0) & (q < (1 << level))).all(dim=1)
        for l in range(level):
            depth = level - l - 1
            child = (((q[:, 0] >> depth) & 1) << 2) | (((q[:, 1] >> depth) & 1) << 1) | ((q[:, 2] >> depth) & 1)
            bits = o[ordn.clamp(max=o.shape[0] - 1)]
            has = ((bits >> child) & 1).bool() & alive
            below = bits & ((2 << child) - 1)
            cnt = torch.zeros_like(below)
            for i in range(8):
                cnt += (below >> i) & 1
            ordn = torch.where(has, psum[ordn.clamp(max=o.shape[0] - 1)] + cnt, ordn)
            alive = has
        return torch.where(alive, ordn - int(self.pyramid[1, level]), torch.full_like(ordn, -1))

    def raytrace(self, ray_o, ray_d, target_level, return_offsets=False):
        """(ray, voxel) nuggets [M,2] int32 at `target_level`, sorted by ray, each ray's run front to back."""
        lib = _lib.load()
        ray_o, ray_d = _f32c(ray_o, "ray_o"), _f32c(ray_d, "ray_d")
        n = ray_o.shape[0]
        dev = ray_o.device
        offsets = torch.empty(n + 1, dtype=torch.int32, device=dev)
        ws = torch.empty(n // 1024 + 2, dtype=torch.int32, device=dev)
        args = (_ptr(self.octree), _ptr(self.prefix), _ptr(self.points), self._pyrsum, self.level, int(target_level),
                _ptr(ray_o), _ptr(ray_d), n)
        with torch.cuda.device(dev):
            _lib.check(lib.nglod_spc_raytrace_count(*args, _ptr(offsets), _ptr(ws), _stream()), "nglod_spc_raytrace_count")
            total = int(offsets[-1])                         # the ONE host read of the traversal
            nuggets = torch.empty(total, 2, dtype=torch.int32, device=dev)
            if total:
                _lib.check(lib.nglod_spc_raytrace_fill(*args, _ptr(offsets), _ptr(nuggets), _stream()),
                           "nglod_spc_raytrace_fill")
        return (nuggets, offsets) if return_offsets else nuggets


def _raytrace_runs(spc, ray_o, ray_d, target_level, capacity=None):
    """One-pass traversal for the in-voxel tracer (nglod_spc_raytrace_runs): per-ray runs in a re-used buffer of
    `capacity` nuggets (default 4 per ray, at least 2^20).  Returns (nuggets, run_begin, run_end, cursor); cursor[1] != 0
    means the buffer was too small (read it AFTER queueing the work that uses the runs, then fall back)."""
    lib = _lib.load()
    n, dev = ray_o.shape[0], ray_o.device
    explicit = capacity is not None
    capacity = max(1 << 20, 4 * n) if capacity is None else int(capacity)
    ws = getattr(spc, "_runs_ws", None)
    if ws is None or ws["n"] != n or ws["dev"] != dev or (ws["cap"] != capacity if explicit else ws["cap"] < capacity):
        ws = {"n": n, "cap": capacity, "dev": dev, "nuggets": torch.empty(capacity, 2, dtype=torch.int32, device=dev),
              "begin": torch.empty(n, dtype=torch.int32, device=dev), "end": torch.empty(n, dtype=torch.int32, device=dev),
              "cursor": torch.zeros(2, dtype=torch.int32, device=dev)}
        spc._runs_ws = ws
    with torch.cuda.device(dev):
        _lib.check(lib.nglod_spc_raytrace_runs(_ptr(spc.octree), _ptr(spc.prefix), _ptr(spc.points), spc._pyrsum, spc.level,
                                               int(target_level), _ptr(ray_o), _ptr(ray_d), n, ws["cap"], _ptr(ws["nuggets"]),
                                               _ptr(ws["begin"]), _ptr(ws["end"]), _ptr(ws["cursor"]), _stream()),
                   "nglod_spc_raytrace_runs")
    return ws["nuggets"], ws["begin"], ws["end"], ws["cursor"]


def mark_first_hit(nuggets):
    lib = _lib.load()
    m = nuggets.shape[0]
    info = torch.empty(m, dtype=torch.int32, device=nuggets.device)
    with torch.cuda.device(nuggets.device):
        _lib.check(lib.nglod_spc_mark_first_hit(_ptr(nuggets), m, _ptr(info), _stream()), "nglod_spc_mark_first_hit")
    return info


def ray_aabb(spc, nuggets, offsets, ray_o, ray_d, level, query=None, active=None, t=None):
    """First occupied voxel per ray from `query` (default: the ray origins).  Returns (x, t, cond, pidx)."""
    lib = _lib.load()
    ray_o, ray_d = _f32c(ray_o, "ray_o"), _f32c(ray_d, "ray_d")
    n = ray_o.shape[0]
    dev = ray_o.device
    query = ray_o if query is None else _f32c(query, "query")
    x = query.clone()
    t = torch.zeros(n, 1, device=dev) if t is None else t.clone()
    cond = torch.zeros(n, dtype=torch.bool, device=dev) if active is None else active.clone()
    pidx = torch.full((n,), -1, dtype=torch.int32, device=dev)
    lp = spc.level_points(level).contiguous()
    with torch.cuda.device(dev):
        _lib.check(lib.nglod_spc_ray_aabb(_ptr(nuggets), _ptr(offsets), n, _ptr(lp), int(level), _ptr(ray_o), _ptr(ray_d),
                                          _ptr(query), _ptr(active), _ptr(x), _ptr(t), _ptr(cond), _ptr(pidx), _stream()),
                   "nglod_spc_ray_aabb")
    return x, t, cond, pidx


# ----------------------------------------------------------------------------------------------- sparse OctreeSDF
class SparseOctreeSDF:
    """An OctreeSDF restricted to the voxels of a sparse octree -- the model the reference's real-time renderer
    traces (sol-renderer/SDF.cu:65-216: corner features `cf`, voxel->8-corner `trinkets` with a `parent` link, one
    35->128->1 decoder per LOD; produced from a trained OctreeSDF by SOL_NGLOD, lib/models/SOL_NGLOD.py:31-100).

    For LOD l the voxels are the octree's level (l + base_lod); every corner of an occupied voxel keeps the dense
    grid's feature vector, so inside occupied voxels `sdf(x, lod, pidx)` equals the dense `OctreeSDF.sdf(x, lod)`.
    Device tables (all LODs concatenated, coarse first):
      corner_feats [NC, F] fp32; trinkets [NV, 8] int32 rows of corner_feats, corner k = bx + 2*by + 4*bz;
      parents [NV] int32 voxel row one LOD up (-1 at LOD 0); voxels [NV, 4] int16; lod_offset[l] = first voxel row.
    """

    def __init__(self, net, spc):
        self.net, self.spc = net, spc
        self.num_lods, self.base_lod = net.num_lods, net.args.base_lod
        if spc.level < self.num_lods + self.base_lod - 1:
            raise ValueError("octree is shallower than the finest LOD")
        dev = spc.octree.device
        feats, feats_summed, trinkets, parents, voxels, lod_offset = [], [], [], [], [], [0]
        # prefix-summed corner rows (nglod_sparse_net_t.corner_feats_summed): the dense model's summed grids sampled at
        # the same corners, so one 8-corner sample of the requested LOD's voxel replaces the walk up the parent chain
        self.sum_lods = bool(getattr(net, "sum_lods", True)) and net._grids_nest() and dev.type == "cuda"
        summed = net._derived_grids()[0] if self.sum_lods else None
        corner_base = 0
        prev_morton = None
        for l in range(self.num_lods):
            level = l + self.base_lod
            S = (1 << level) + 1
            vox = spc.level_points(level)[:, :3].long()                       # Morton order
            off = torch.tensor([[k & 1, (k >> 1) & 1, (k >> 2) & 1] for k in range(8)], device=dev)
            cor = vox.unsqueeze(1) + off.unsqueeze(0)                         # [nv, 8, 3]
            key = (cor[..., 2] * S + cor[..., 1]) * S + cor[..., 0]
            uniq, inv = torch.unique(key.reshape(-1), return_inverse=True)
            cz, cy, cx = uniq // (S * S), (uniq // S) % S, uniq % S
            fm = net.features[l].fm.data                                      # [1, F, D, H, W]
            feats.append(fm[0][:, cz, cy, cx].t().contiguous())
            if summed is not None:
                feats_summed.append(summed[l][0][:, cz, cy, cx].t().contiguous())
            trinkets.append((inv.reshape(-1, 8) + corner_base).int())
            morton = points_to_morton(vox)
            if l == 0:
                parents.append(torch.full((vox.shape[0],), -1, dtype=torch.int32, device=dev))
            else:
                pi = torch.searchsorted(prev_morton, morton >> 3)
                parents.append((pi + lod_offset[l - 1]).int())
            prev_morton = morton
            v4 = torch.zeros(vox.shape[0], 4, dtype=torch.int16, device=dev)
            v4[:, :3] = vox.short()
            voxels.append(v4)
            corner_base += uniq.shape[0]
            lod_offset.append(lod_offset[-1] + vox.shape[0])
        self.corner_feats = torch.cat(feats).contiguous()
        self.corner_feats_summed = torch.cat(feats_summed).contiguous() if summed is not None else None
        self.trinkets = torch.cat(trinkets).contiguous()
        self.parents = torch.cat(parents).contiguous()
        self.voxels = torch.cat(voxels).contiguous()
        self.lod_offset = lod_offset
        self.math_mode = getattr(net, "math_mode", "tc")

    def _decoder_params(self, lod):
        if self.net is None:
            return self._decoders[lod]
        return self.net.decoder_params(lod)

    @classmethod
    def load(cls, path, device="cuda", math_mode="tc"):
        """Read the real-time renderer's model file back (what sol-renderer/SDF.cu:65-139 `loadWeights` + :141-216
        `initTrinkets` do): `octree` bytes -> SPC; `cc` (uint8 corner coordinates, all LODs concatenated, `pyramid[l]`
        rows each) + `cf` (fp16 corner features) -> corner table; voxel -> 8 corner rows (`trinkets`) and the parent
        link are re-derived from the coordinates -- the reference searches `cc` linearly per corner (index_trinket.cuh),
        here: sort the corner keys once, binary-search every voxel corner.  Decoders come back as fp32 copies of the
        stored fp16 values.  The file format fixes base_lod = 2 (LOD l <-> octree level l + 2, SDF.cu:155-158)."""
        import numpy as np
        z = np.load(path)
        dev = torch.device(device)
        self = cls.__new__(cls)
        self.net, self.math_mode = None, math_mode
        self.spc = SPC(torch.from_numpy(z["octree"].astype(np.uint8)).to(dev))
        counts = [int(c) for c in z["pyramid"]]
        self.num_lods, self.base_lod = len(counts), 2
        if self.spc.level < self.num_lods + 1:
            raise ValueError("octree is shallower than the file's finest LOD")
        cc = torch.from_numpy(z["cc"].astype(np.int64)).to(dev)
        if cc.shape[0] != sum(counts) or z["cf"].shape[0] != sum(counts):
            raise ValueError("cc / cf / pyramid disagree on the number of corner rows")
        self.corner_feats = torch.from_numpy(z["cf"].astype(np.float32)).to(dev).contiguous()
        off = torch.tensor([[k & 1, (k >> 1) & 1, (k >> 2) & 1] for k in range(8)], device=dev)
        trinkets, parents, voxels, lod_offset = [], [], [], [0]
        base, prev_morton = 0, None
        for l, cnt in enumerate(counts):
            level = l + 2
            S = (1 << level) + 1
            c = cc[base:base + cnt]
            ckey = (c[:, 2] * S + c[:, 1]) * S + c[:, 0]
            skey, order = torch.sort(ckey)
            vox = self.spc.level_points(level)[:, :3].long()
            cor = vox.unsqueeze(1) + off.unsqueeze(0)
            vkey = ((cor[..., 2] * S + cor[..., 1]) * S + cor[..., 0]).reshape(-1)
            pos = torch.searchsorted(skey, vkey).clamp(max=cnt - 1)
            if not bool((skey[pos] == vkey).all()):
                raise ValueError(f"LOD {l}: a voxel corner is missing from cc (file and octree disagree)")
            trinkets.append((order[pos].reshape(-1, 8) + base).int())
            morton = points_to_morton(vox)
            if l == 0:
                parents.append(torch.full((vox.shape[0],), -1, dtype=torch.int32, device=dev))
            else:
                parents.append((torch.searchsorted(prev_morton, morton >> 3) + lod_offset[l - 1]).int())
            prev_morton = morton
            v4 = torch.zeros(vox.shape[0], 4, dtype=torch.int16, device=dev)
            v4[:, :3] = vox.short()
            voxels.append(v4)
            base += cnt
            lod_offset.append(lod_offset[-1] + vox.shape[0])
        self.trinkets = torch.cat(trinkets).contiguous()
        self.parents = torch.cat(parents).contiguous()
        self.voxels = torch.cat(voxels).contiguous()
        self.lod_offset = lod_offset
        f32 = lambda a: torch.from_numpy(a.astype(np.float32)).to(dev).contiguous()
        self._decoders = [(f32(z["w0"][i]), f32(z["b0"][i]), f32(z["w1"][i]), f32(z["b1"][i])) for i in range(self.num_lods)]
        self.pos_invariant = self._decoders[0][0].shape[1] == self.corner_feats.shape[1]
        self.sum_lods = dev.type == "cuda"
        self.corner_feats_summed = summed_corner_rows(self.corner_feats, self.trinkets, self.parents, self.voxels,
                                                      self.lod_offset, self.num_lods) if self.sum_lods else None
        return self

    def save(self, path):
        """Write the reference's real-time renderer format (SOL_NGLOD.save, lib/models/SOL_NGLOD.py:80-100; read by
        sol-renderer/SDF.cu:65-139): octree bytes, corner coordinates `cc` (uint8), corner features `cf` and the
        decoders in fp16, and `pyramid` = corner rows per LOD."""
        import numpy as np
        cc, counts = [], []
        row = 0
        for l in range(self.num_lods):
            v = slice(self.lod_offset[l], self.lod_offset[l + 1])
            tr = self.trinkets[v].long()
            nrows = int(tr.max()) + 1 - row
            coords = torch.zeros(nrows, 3, dtype=torch.uint8, device=tr.device)
            vox = self.voxels[v, :3].long()
            for k in range(8):
                off = torch.tensor([k & 1, (k >> 1) & 1, (k >> 2) & 1], device=tr.device)
                coords[tr[:, k] - row] = (vox + off).to(torch.uint8)
            cc.append(coords)
            counts.append(nrows)
            row += nrows
        dec = [self._decoder_params(i) for i in range(self.num_lods)]
        np.savez_compressed(
            path, octree=self.spc.octree.cpu().numpy(), cc=torch.cat(cc).cpu().numpy(),
            cf=self.corner_feats.half().cpu().numpy(),
            w0=torch.stack([d[0].detach().half() for d in dec]).cpu().numpy(),
            b0=torch.stack([d[1].detach().half() for d in dec]).cpu().numpy(),
            w1=torch.stack([d[2].detach().half() for d in dec]).cpu().numpy(),
            b1=torch.stack([d[3].detach().half() for d in dec]).cpu().numpy(),
            pyramid=np.array(counts))

    def struct(self):
        s = _lib.SparseNetStruct()
        s.num_lods, s.base_lod = self.num_lods, self.base_lod
        s.feature_dim, s.hidden_dim = self.corner_feats.shape[1], self._decoder_params(0)[0].shape[0]
        s.math_mode = _lib.MATH_TC3XTF32 if self.math_mode == "tc" else _lib.MATH_FP32
        s.pos_invariant = 1 if getattr(self, "pos_invariant", False) else 0
        s.corner_feats, s.trinkets = self.corner_feats.data_ptr(), self.trinkets.data_ptr()
        s.parents, s.voxels = self.parents.data_ptr(), self.voxels.data_ptr()
        if self.sum_lods and getattr(self, "corner_feats_summed", None) is not None:
            s.corner_feats_summed = self.corner_feats_summed.data_ptr()
        for i, o in enumerate(self.lod_offset):
            s.lod_voxel_offset[i] = o
        self._keep = [tuple(p.data for p in self._decoder_params(i)) for i in range(self.num_lods)]
        for i, (w0, b0, w1, b1) in enumerate(self._keep):
            s.w0[i], s.b0[i], s.w1[i], s.b1[i] = w0.data_ptr(), b0.data_ptr(), w1.data_ptr(), b1.data_ptr()
        return s

    def sdf(self, x, lod, pidx):
        """sdf at x inside (or near) voxel `pidx` (index within level lod + base_lod): [N,3], [N] -> [N,1]."""
        lib = _lib.load()
        x = _f32c(x, "x")
        pidx = pidx.int().contiguous()
        n = x.shape[0]
        out = torch.empty(n, 1, device=x.device)
        s = self.struct()
        with torch.cuda.device(x.device):
            _lib.check(lib.nglod_sparse_sdf_forward(ctypes.byref(s), int(lod), _ptr(x), _ptr(pidx), n, _ptr(out), _stream()),
                       "nglod_sparse_sdf_forward")
        return out

    def trace(self, ray_o, ray_d, lod, num_steps=50, min_dis=0.0003, far=5.0, normal_h=0.001, stats=None, one_pass=True,
              runs_capacity=None):
        """The reference renderer's frame: traverse -> first voxel -> in-voxel sphere trace with re-location
        (sol-renderer/sdfRenderer.cu:176-260 + SDF.cu:297-472).  Returns (x, depth, hit, normal, pidx).
        one_pass (default): the traversal writes per-ray runs in one pass (nglod_spc_raytrace_runs) and the tracer starts
        without a host read in between; the "buffer too small" flag is read after the frame has been queued, and the frame
        is redone through the exact two-pass list (`spc.raytrace`, the reference's nugget order) if it was set.
        one_pass=False: always the two-pass list.  Same result either way (the tracer only ever walks a ray's own run)."""
        lib = _lib.load()
        ray_o, ray_d = _f32c(ray_o, "ray_o"), _f32c(ray_d, "ray_d")
        n, dev = ray_o.shape[0], ray_o.device
        cursor = None
        if one_pass and n > 0:
            nuggets, run_begin, run_end, cursor = _raytrace_runs(self.spc, ray_o, ray_d, lod + self.base_lod, runs_capacity)
        else:
            nuggets, offsets = self.spc.raytrace(ray_o, ray_d, lod + self.base_lod, return_offsets=True)
            run_begin, run_end = offsets[:-1], offsets[1:]
        x = torch.empty(n, 3, device=dev)
        depth = torch.empty(n, 1, device=dev)
        hit = torch.empty(n, dtype=torch.bool, device=dev)
        normal = torch.empty(n, 3, device=dev)
        pidx = torch.empty(n, dtype=torch.int32, device=dev)
        queue = torch.empty(1, dtype=torch.int32, device=dev)
        opts = _lib.TraceOpts(int(num_steps), 1, 1.0, float(min_dis), float(far), float(normal_h))
        s = self.struct()
        with torch.cuda.device(dev):
            _lib.check(lib.nglod_spc_sphere_trace_runs(ctypes.byref(s), int(lod), _ptr(nuggets), _ptr(run_begin), _ptr(run_end),
                                                       _ptr(ray_o), _ptr(ray_d), n, ctypes.byref(opts), _ptr(x), _ptr(depth),
                                                       _ptr(hit), _ptr(normal), _ptr(pidx), _ptr(queue), _ptr(stats), _stream()),
                       "nglod_spc_sphere_trace_runs")
        if cursor is not None and int(cursor[1]) != 0:       # read after everything is queued: no bubble in the common case
            if stats is not None:
                stats.zero_()
            return self.trace(ray_o, ray_d, lod, num_steps, min_dis, far, normal_h, stats, one_pass=False)
        return x, depth, hit, normal, pidx


# ------------------------------------------------------------------------------------------------ sparse training
def build_sparse_tables(spc, num_lods, base_lod):
    """The reference's create_dual / create_trinkets (app/spc/spc_utils.py:91-150, a CPU dict + pandas join) as device
    ops: per LOD l (octree level l + base_lod) the unique corners of the occupied voxels (sort + unique), every voxel's 8
    corner rows (`trinkets`, corner k = bx + 2 by + 4 bz), its parent voxel one LOD up (binary search on Morton codes)
    and its integer coordinates.  Returns (corner_counts, trinkets [NV,8] i32, parents [NV] i32, voxels [NV,4] i16,
    lod_offset, corner_xyz [NC,3] long)."""
    dev = spc.octree.device
    trinkets, parents, voxels, counts, cxyz, lod_offset = [], [], [], [], [], [0]
    corner_base, prev_morton = 0, None
    off = torch.tensor([[k & 1, (k >> 1) & 1, (k >> 2) & 1] for k in range(8)], device=dev)
    for l in range(num_lods):
        level = l + base_lod
        S = (1 << level) + 1
        vox = spc.level_points(level)[:, :3].long()
        cor = vox.unsqueeze(1) + off.unsqueeze(0)
        key = (cor[..., 2] * S + cor[..., 1]) * S + cor[..., 0]
        uniq, inv = torch.unique(key.reshape(-1), return_inverse=True)
        cxyz.append(torch.stack([uniq % S, (uniq // S) % S, uniq // (S * S)], dim=1))
        trinkets.append((inv.reshape(-1, 8) + corner_base).int())
        morton = points_to_morton(vox)
        if l == 0:
            parents.append(torch.full((vox.shape[0],), -1, dtype=torch.int32, device=dev))
        else:
            parents.append((torch.searchsorted(prev_morton, morton >> 3) + lod_offset[l - 1]).int())
        prev_morton = morton
        v4 = torch.zeros(vox.shape[0], 4, dtype=torch.int16, device=dev)
        v4[:, :3] = vox.short()
        voxels.append(v4)
        counts.append(int(uniq.shape[0]))
        corner_base += int(uniq.shape[0])
        lod_offset.append(lod_offset[-1] + vox.shape[0])
    return (counts, torch.cat(trinkets).contiguous(), torch.cat(parents).contiguous(), torch.cat(voxels).contiguous(),
            lod_offset, torch.cat(cxyz).contiguous())


def summed_corner_rows(cf, trinkets, parents, voxels, lod_offset, num_lods):
    """Prefix-summed corner rows (nglod_sparse_net_t.corner_feats_summed): row of a LOD-l corner = sum over k <= l of the
    level-k interpolant at that corner, built level by level (each corner is evaluated in the parent of a voxel that
    owns it; voxels sharing a corner agree because the coarser field is continuous across occupied voxels)."""
    out = cf.clone()
    dev = cf.device
    off = torch.tensor([[k & 1, (k >> 1) & 1, (k >> 2) & 1] for k in range(8)], device=dev)
    for l in range(1, num_lods):
        v0, v1 = lod_offset[l], lod_offset[l + 1]
        tr = trinkets[v0:v1].long()
        par = parents[v0:v1].long()
        vox = voxels[v0:v1, :3].long()
        pvox = voxels[par, :3].long()
        ptr = trinkets[par].long()                                    # parent's 8 corner rows (already summed)
        pval = out[ptr]                                               # [nv, 8, F]
        for k in range(8):
            f = ((vox + off[k]) - 2 * pvox).float() * 0.5             # position of corner k in the parent, in {0, .5, 1}^3
            g = 1.0 - f
            acc = 0
            for j in range(8):
                w = (f[:, 0] if j & 1 else g[:, 0]) * (f[:, 1] if j & 2 else g[:, 1]) * (f[:, 2] if j & 4 else g[:, 2])
                acc = acc + w.unsqueeze(1) * pval[:, j]
            out[tr[:, k]] = cf[tr[:, k]] + acc
    return out


class _SparseSdfFunction(torch.autograd.Function):
    """d = NeuralSPC.sdf(x, lod, pidx): sparse forward kernel; the backward recomputes it in-kernel and scatters
    dL/d(corner features) along each query's parent chain (nglod_sparse_sdf_backward)."""

    @staticmethod
    def forward(ctx, x, pidx, module, lod, corner_feats, w0, b0, w1, b1):
        ctx.module, ctx.lod = module, lod
        ctx.save_for_backward(x, pidx)
        return SparseOctreeSDF.sdf(module, x, lod, pidx)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        x, pidx = ctx.saved_tensors
        m, lod = ctx.module, ctx.lod
        lib = _lib.load()
        needs = ctx.needs_input_grad
        g_cf = torch.zeros_like(m.corner_feats) if needs[4] else None
        gd = [torch.zeros_like(p) if needs[5 + k] else None for k, p in enumerate(m._decoder_params(lod))]
        s = m.struct()
        xx = _f32c(x, "x")
        go = _f32c(grad_out, "grad_out").reshape(-1)
        with torch.cuda.device(xx.device):
            _lib.check(lib.nglod_sparse_sdf_backward(ctypes.byref(s), int(lod), _ptr(xx), _ptr(pidx.int().contiguous()),
                                                     xx.shape[0], _ptr(go), _ptr(g_cf), _ptr(gd[0]), _ptr(gd[1]),
                                                     _ptr(gd[2]), _ptr(gd[3]), _stream()), "nglod_sparse_sdf_backward")
        return (None, None, None, None, g_cf, *gd)


class NeuralSPC(torch.nn.Module, SparseOctreeSDF):
    """A natively sparse, trainable OctreeSDF -- the model of the reference's app/spc (NeuralSPC.py:39-145, SPC.py:35-107):
    features live ONLY on the corners of the occupied voxels of an octree (`corner_feats`, one nn.Parameter, all LODs
    concatenated coarse first, N(0, feature_std) init), one decoder per LOD.  `sdf(x, lod, pidx)` sums the trilinear
    samples along the voxel's parent chain and decodes; it is differentiable (parameters only) through the sparse
    kernels.  `pos_invariant=True` gives the reference's feature-only decoders (BasicDecoder(feature_dim -> 1));
    the default concatenates [x, features] like OctreeSDF.  The level-7+ models this enables have no dense grid at all."""

    def __init__(self, spc, num_lods, base_lod=2, feature_dim=32, hidden_dim=128, feature_std=0.01, pos_invariant=False,
                 math_mode="tc"):
        torch.nn.Module.__init__(self)
        if spc.level < num_lods + base_lod - 1:
            raise ValueError("octree is shallower than the finest LOD")
        self.spc, self.num_lods, self.base_lod = spc, num_lods, base_lod
        self.pos_invariant, self.math_mode, self.sum_lods = pos_invariant, math_mode, True
        self.corner_feats_summed = None
        counts, self.trinkets, self.parents, self.voxels, self.lod_offset, self.corner_xyz = \
            build_sparse_tables(spc, num_lods, base_lod)
        self.corner_counts = counts
        dev = spc.octree.device
        self.corner_feats = torch.nn.Parameter(torch.randn(sum(counts), feature_dim, device=dev) * feature_std)
        in_dim = feature_dim + (0 if pos_invariant else 3)
        self.louts = torch.nn.ModuleList([
            torch.nn.Sequential(torch.nn.Linear(in_dim, hidden_dim), torch.nn.ReLU(), torch.nn.Linear(hidden_dim, 1))
            for _ in range(num_lods)]).to(dev)
        self.lod = None

    def _decoder_params(self, lod):
        seq = self.louts[lod]
        return (seq[0].weight, seq[0].bias, seq[2].weight, seq[2].bias)

    def summed_rows(self):
        """Prefix-summed corner rows (nglod_sparse_net_t.corner_feats_summed), see summed_corner_rows."""
        return summed_corner_rows(self.corner_feats.data, self.trinkets, self.parents, self.voxels, self.lod_offset, self.num_lods)

    def struct(self):
        # inference (eval mode, no autograd): sample ONE level from the prefix-summed rows, rebuilt when the features change
        if self.sum_lods and not self.training and not torch.is_grad_enabled():
            key = (self.corner_feats._version, self.corner_feats.data_ptr())
            if getattr(self, "_summed_key", None) != key:
                self.corner_feats_summed, self._summed_key = self.summed_rows(), key
        else:
            self.corner_feats_summed = None
            self._summed_key = None
        return SparseOctreeSDF.struct(self)

    def query(self, x, lod):
        """Voxel index (within LOD `lod`'s level) containing each point, -1 outside the octree (SPC.query, SPC.py:86-90)."""
        level = lod + self.base_lod
        return self.spc.query(quantize_points(x, level), level)

    def sdf(self, x, lod=None, pidx=None):
        lod = (self.num_lods - 1 if self.lod is None else self.lod) if lod is None else lod
        if pidx is None:
            pidx = self.query(x, lod)
        params = self._decoder_params(lod)
        if torch.is_grad_enabled() and (self.corner_feats.requires_grad or any(p.requires_grad for p in params)):
            return _SparseSdfFunction.apply(x, pidx, self, lod, self.corner_feats, *params)
        return SparseOctreeSDF.sdf(self, x, lod, pidx)

    def loss_backward(self, x, gt, lods=None, pidx=None, global_batch=None):
        """Fused training step without the optimiser: for every head in `lods` (default: all) ONE kernel evaluates
        d = sdf(x, lod), adds sum((d - gt)^2) / global_batch to the loss and accumulates dL/d(corner_feats) and
        dL/d(decoder of that head) into the parameters' `.grad` (nglod_sparse_sdf_train_step) -- what
        `sum(((net.sdf(x, l) - gt) ** 2).sum() for l in lods) / B` followed by `.backward()` computes through autograd,
        minus the separate forward launch, the loss kernels and the saved tensors.  Points outside the octree (pidx < 0) are
        inert rows: they add no loss and no gradient (and `sdf` returns 0 for them).  Returns the per-head losses [len(lods)]."""
        lods = list(range(self.num_lods)) if lods is None else list(lods)
        lib = _lib.load()
        xx = _f32c(x, "x")
        g = _f32c(gt, "gt").reshape(-1)
        n = xx.shape[0]
        scale = 1.0 / float(global_batch if global_batch is not None else max(n, 1))
        if self.corner_feats.grad is None:
            self.corner_feats.grad = torch.zeros_like(self.corner_feats)
        losses = torch.zeros(len(lods), device=xx.device, dtype=torch.float32)
        s = self.struct()
        with torch.cuda.device(xx.device):
            for k, lod in enumerate(lods):
                pi = (self.query(xx, lod) if pidx is None else pidx[k]).int().contiguous()
                params = self._decoder_params(lod)
                for p in params:
                    if p.grad is None:
                        p.grad = torch.zeros_like(p)
                _lib.check(lib.nglod_sparse_sdf_train_step(ctypes.byref(s), int(lod), _ptr(xx), _ptr(pi), _ptr(g), n, scale,
                                                           _ptr(self.corner_feats.grad), _ptr(params[0].grad),
                                                           _ptr(params[1].grad), _ptr(params[2].grad), _ptr(params[3].grad),
                                                           _ptr(losses[k:k + 1]), _stream()), "nglod_sparse_sdf_train_step")
        return losses

    def forward(self, x):
        return self.sdf(x)
