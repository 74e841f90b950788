"""CPU oracle for the NGLOD hot path (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

A restatement of the reference's PyTorch path with the same ATen calls the reference makes
(F.grid_sample / F.linear / torch.where / addcmul), so it runs multi-threaded on the host and
is the "port" CPU baseline, plus an independent float64 numpy restatement of the interpolation
arithmetic that cross-checks it.

Parity status: PINNED.  tests/golden/make_golden.py imports the unmodified reference from
/root/reference in the authoring container and stores its outputs; tests/test_oracle_golden.py
checks this file against those vectors (sdf forward, gradients, SphereTracer.forward, Renderer.render).
The native pieces (aabb, mesh2sdf) live in oracle.c and are pinned on the GPU against oracle/_ref.

Each function cites the reference lines it follows (paths relative to the nv-tlabs/nglod tree).
"""
import ctypes
import os
import subprocess

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBO = os.path.join(_HERE, "liboracle.so")


# --------------------------------------------------------------------------- C part
def build_c(force=False):
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIBO) or os.path.getmtime(_LIBO) < os.path.getmtime(src):
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", src, "-o", _LIBO, "-lm"], check=True)
    return _LIBO


_c = None


def _clib():
    global _c
    if _c is None:
        lib = ctypes.CDLL(build_c())
        vp, i64 = ctypes.c_void_p, ctypes.c_int64
        lib.oracle_aabb.argtypes = [vp, vp, i64, vp, vp, vp]
        lib.oracle_aabb.restype = None
        lib.oracle_mesh2sdf.argtypes = [vp, i64, vp, i64, vp]
        lib.oracle_mesh2sdf.restype = None
        _c = lib
    return _c


def aabb(ray_o, ray_d):
    """sol_nglod.aabb on the CPU: returns (x [N,3], t [N,1], hit [N] bool).  oracle.c:oracle_aabb."""
    o = ray_o.detach().cpu().float().contiguous()
    d = ray_d.detach().cpu().float().contiguous()
    n = o.shape[0]
    x = torch.empty(n, 3)
    t = torch.empty(n, 1)
    hit = torch.empty(n, dtype=torch.uint8)
    _clib().oracle_aabb(o.data_ptr(), d.data_ptr(), n, x.data_ptr(), t.data_ptr(), hit.data_ptr())
    return x, t, hit.bool()


def mesh2sdf(points, mesh):
    """mesh2sdf.mesh2sdf_gpu on the CPU (O(N*T), keep N*T small): returns [N] signed distance."""
    p = points.detach().cpu().float().contiguous()
    m = mesh.detach().cpu().float().contiguous()
    out = torch.empty(p.shape[0])
    _clib().oracle_mesh2sdf(p.data_ptr(), p.shape[0], m.data_ptr(), m.shape[0], out.data_ptr())
    return out


# --------------------------------------------------------------------------- the model
class OracleNet:
    """OctreeSDF parameters in the REFERENCE layout (fm [1,C,D,H,W] contiguous), built from a
    state_dict with the reference's keys (features.{i}.fm, louts.{i}.{0,2}.{weight,bias})."""

    def __init__(self, state_dict, pos_invariant=False, device="cpu", requires_grad=False):
        sd = {k: v.detach().to(device=device, dtype=torch.float32) for k, v in state_dict.items()}
        self.num_lods = len([k for k in sd if k.startswith("features.") and k.endswith(".fm")])
        self.fm = [sd[f"features.{i}.fm"].contiguous().clone() for i in range(self.num_lods)]
        n_dec = len([k for k in sd if k.startswith("louts.") and k.endswith(".0.weight")])
        self.dec = [tuple(sd[f"louts.{i}.{k}"].contiguous().clone() for k in ("0.weight", "0.bias", "2.weight", "2.bias"))
                    for i in range(n_dec)]
        self.pos_invariant = pos_invariant
        self.lod = None
        if requires_grad:
            for t in self.parameters():
                t.requires_grad_(True)

    def parameters(self):
        return list(self.fm) + [t for d in self.dec for t in d]

    def decoder(self, lod):
        return self.dec[0 if len(self.dec) == 1 else lod]

    # FeatureVolume.forward, OctreeSDF.py:46-57
    def sample(self, i, x):
        n = x.shape[0]
        grid = x.reshape(1, n, 1, 1, 3)
        return F.grid_sample(self.fm[i], grid, align_corners=True, padding_mode="border")[0, :, :, 0, 0].transpose(0, 1)

    # OctreeSDF.sdf, OctreeSDF.py:94-155 (integer lod, or all heads)
    def sdf(self, x, lod=None, return_lst=False):
        if lod is None:
            lod = self.lod
        preds, feat = [], None
        for i in range(self.num_lods):
            s = self.sample(i, x)
            feat = s if feat is None else s + feat                      # :109-110 running sum
            inp = feat if self.pos_invariant else torch.cat([x, feat], dim=-1)   # :113-115
            w0, b0, w1, b1 = self.decoder(i)
            d = F.linear(F.relu(F.linear(inp, w0, b0)), w1, b1)          # :84-86,142
            if lod is not None and lod == i:
                return d
            preds.append(d)
        return preds if return_lst else preds[-1]

    def __call__(self, x):                                               # BaseLOD.forward, BaseLOD.py:37-41
        return self.sdf(x)


def sdf_explicit_f64(net, x, lod):
    """Independent restatement of sdf(x, lod) in float64 numpy with the interpolation written out
    (PyTorch grid_sampler_3d: unnormalise with align_corners, clip to the border, floor, 8 weights)."""
    p = x.detach().cpu().double().numpy()
    feat = np.zeros((p.shape[0], net.fm[0].shape[1]))
    for i in range(lod + 1):
        fm = net.fm[i].detach().cpu().double().numpy()[0]               # [C, D(z), H(y), W(x)]
        R = fm.shape[-1] - 1
        u = np.clip((p + 1.0) / 2.0 * R, 0.0, R)
        i0 = np.floor(u).astype(np.int64)
        f = u - i0
        i1 = np.minimum(i0 + 1, R)
        for bz in (0, 1):
            for by in (0, 1):
                for bx in (0, 1):
                    ix = i1[:, 0] if bx else i0[:, 0]
                    iy = i1[:, 1] if by else i0[:, 1]
                    iz = i1[:, 2] if bz else i0[:, 2]
                    w = (f[:, 0] if bx else 1 - f[:, 0]) * (f[:, 1] if by else 1 - f[:, 1]) * (f[:, 2] if bz else 1 - f[:, 2])
                    feat += w[:, None] * fm[:, iz, iy, ix].T
    w0, b0, w1, b1 = (t.detach().cpu().double().numpy() for t in net.decoder(lod))
    inp = feat if net.pos_invariant else np.concatenate([p, feat], axis=1)
    h = np.maximum(inp @ w0.T + b0, 0.0)
    return h @ w1.T + b1


# --------------------------------------------------------------------------- gradients / tracer
def gradient_finitediff(x, f):
    """diffutils.gradient(..., 'finitediff'), diffutils.py:61-70."""
    h = 1.0 / (64.0 * 3.0)
    ex = torch.tensor([h, 0.0, 0.0], device=x.device)
    ey = torch.tensor([0.0, h, 0.0], device=x.device)
    ez = torch.tensor([0.0, 0.0, h], device=x.device)
    g = torch.cat([f(x + ex) - f(x - ex), f(x + ey) - f(x - ey), f(x + ez) - f(x - ez)], dim=-1)
    return g / (h * 2.0)


def sphere_trace(net, ray_o, ray_d, num_steps=256, step_size=1.0, min_dis=0.0003, far=10.0, aabb_fn=None,
                 count=None):
    """SphereTracer.forward, SphereTracer.py:41-132, restated line for line as the batch loop.
    Returns dict(x, depth, hit, normal).  `count`, if a dict, receives the number of SDF queries."""
    aabb_fn = aabb_fn or aabb
    x, t, cond = aabb_fn(ray_o, ray_d)                                   # :53
    x, t, cond = x.to(ray_o.device), t.to(ray_o.device), cond.to(ray_o.device)
    normal = torch.zeros_like(x)
    nq = 0
    with torch.no_grad():
        d = net(x)                                                       # :64
        nq += x.shape[0]
        dprev = d.clone()
        hit = torch.zeros_like(d).byte()
        for _ in range(num_steps):                                       # :74
            hit = (torch.abs(t) < far)[:, 0]                             # :84
            cond = cond & (torch.abs(d) > min_dis)[:, 0]                 # :87
            cond = cond & (torch.abs((d + dprev) / 2.0) > min_dis * 3)[:, 0]   # :90
            cond = cond & hit                                            # :93
            if not cond.any():                                           # :98
                break
            x = torch.where(cond.view(-1, 1), torch.addcmul(ray_o, ray_d, t), x)   # :102
            dprev = torch.where(cond.unsqueeze(1), d, dprev)             # :105
            d[cond] = net(x[cond]) * step_size                           # :109
            nq += int(cond.sum())
            t = torch.where(cond.view(-1, 1), t + d, t)                  # :114
        hit = hit.bool() & ~(torch.abs(x) > 1.0).any(dim=-1)             # :119
        g = gradient_finitediff(x[hit], net)                             # :128
        nq += 6 * int(hit.sum())
        normal[hit] = F.normalize(g, p=2, dim=-1, eps=1e-5)              # :129-130
    if count is not None:
        count["sdf_queries"] = nq
    return dict(x=x, depth=t, hit=hit, normal=normal)


def look_at(f, t, width, height, mode="persp", fov=30.0, device="cpu"):
    """geoutils.look_at + normalized_grid, geoutils.py:140-154,180-206 (jitter from torch.rand)."""
    origin = torch.tensor(list(f), dtype=torch.float32, device=device)
    view = F.normalize(torch.tensor(list(t), dtype=torch.float32, device=device) - origin, dim=0)
    right = F.normalize(torch.linalg.cross(view, torch.tensor([0.0, 1.0, 0.0], device=device)), dim=0)
    up = F.normalize(torch.linalg.cross(right, view), dim=0)
    wx = torch.linspace(-1, 1, steps=width, device=device) * (width / height)
    wx += torch.rand(*wx.shape, device=device) * (1.0 / width)
    wy = torch.linspace(1, -1, steps=height, device=device)
    wy += torch.rand(*wy.shape, device=device) * (1.0 / height)
    coord = torch.stack(torch.meshgrid(wx, wy, indexing="ij")).permute(1, 2, 0)
    tan = np.tan(np.radians(fov / 2))
    plane = (right * coord[..., 0, None] * tan + up * coord[..., 1, None] * tan + origin + view).reshape(-1, 3)
    if mode == "ortho":
        return plane, F.normalize(view.unsqueeze(0).repeat(plane.shape[0], 1), dim=-1)
    ray_d = F.normalize(plane - origin, dim=-1)
    return origin.repeat(ray_d.shape[0], 1), ray_d


def l2_loss_and_grads(net, x, gt, lods):
    """Trainer.step_geometry's objective, trainer.py:317-339: sum_l sum_i (sdf_l(x_i)-gt_i)^2 / batch,
    differentiated with torch autograd.  `net` must have been built with requires_grad=True."""
    for p in net.parameters():
        p.grad = None
    loss = 0
    for l in lods:
        loss = loss + ((net.sdf(x, lod=l) - gt) ** 2).sum()
    loss = loss / x.shape[0]
    loss.backward()
    return loss.detach()


# --------------------------------------------------------------------------- SPC (sparse octree)
def spc_raytrace(octree, prefix, points, pyramid, target, ray_o, ray_d):
    """CPU restatement of the reference's level-synchronous traversal (oracle.c:oracle_spc_raytrace).
    Returns nuggets [M,2] int32 and per-ray counts."""
    lib = _clib()
    lib.oracle_spc_raytrace.restype = ctypes.c_int64
    o = octree.cpu().contiguous()
    pf = prefix.cpu().int().contiguous()
    pts = points.cpu().short().contiguous()
    ps = pyramid[1].cpu().int().contiguous()
    ro, rd = ray_o.cpu().float().contiguous(), ray_d.cpu().float().contiguous()
    n = ro.shape[0]
    counts = torch.zeros(n, dtype=torch.int32)
    vp = ctypes.c_void_p
    args = [vp(o.data_ptr()), vp(pf.data_ptr()), vp(pts.data_ptr()), vp(ps.data_ptr()), ctypes.c_int(int(target)),
            vp(ro.data_ptr()), vp(rd.data_ptr()), ctypes.c_int64(n)]
    total = lib.oracle_spc_raytrace(*args, vp(0), ctypes.c_int64(0), vp(counts.data_ptr()))
    nug = torch.zeros(total, 2, dtype=torch.int32)
    lib.oracle_spc_raytrace(*args, vp(nug.data_ptr()), ctypes.c_int64(total), vp(counts.data_ptr()))
    return nug, counts


def spc_ray_aabb(nuggets, level_points, level, ray_o, ray_d, query=None, active=None, x=None, t=None, cond=None, pidx=None):
    """oracle.c:oracle_spc_ray_aabb; with `active` (bool [n]) only those rays are re-located and x/t/cond/pidx are
    updated in place from the given state.  Returns (x, t, cond, pidx)."""
    lib = _clib()
    lib.oracle_spc_ray_aabb.restype = None
    nug = nuggets.cpu().int().contiguous()
    lp = level_points.cpu().short().contiguous()
    ro, rd = ray_o.cpu().float().contiguous(), ray_d.cpu().float().contiguous()
    q = ro if query is None else query.cpu().float().contiguous()
    n = ro.shape[0]
    x = q.clone() if x is None else x.float().contiguous().clone()
    t = torch.zeros(n, 1) if t is None else t.float().contiguous().clone()
    cond8 = torch.zeros(n, dtype=torch.uint8) if cond is None else cond.to(torch.uint8).contiguous().clone()
    pidx = torch.full((n,), -1, dtype=torch.int32) if pidx is None else pidx.int().contiguous().clone()
    act = None if active is None else active.to(torch.uint8).contiguous()
    vp = ctypes.c_void_p
    lib.oracle_spc_ray_aabb(vp(nug.data_ptr()), ctypes.c_int64(nug.shape[0]), vp(lp.data_ptr()), ctypes.c_int(int(level)),
                            vp(ro.data_ptr()), vp(rd.data_ptr()), vp(q.data_ptr()), vp(act.data_ptr() if act is not None else 0),
                            vp(x.data_ptr()), vp(t.data_ptr()), vp(cond8.data_ptr()), vp(pidx.data_ptr()))
    return x, t, cond8.bool(), pidx


class OracleSparseNet:
    """CPU twin of the sparse OctreeSDF tables (nglod_b200.lib.spc.SparseOctreeSDF): corner features, trinkets,
    parents, voxel coordinates, decoders.  Follows sol-renderer/include/solr/solr/sdf/sparse_grid_sample.cuh:31-109
    (parent-chain walk, un-clamped trilinear weights from the voxel's own coordinates) and SDF.cu:412-413."""

    def __init__(self, corner_feats, trinkets, parents, voxels, lod_offset, base_lod, decoders, pos_invariant=False):
        self.pos_invariant = pos_invariant           # feature-only decoders (the reference's NeuralSPC.py:91-95)
        self.cf = corner_feats.detach().cpu().float()
        self.trinkets = trinkets.detach().cpu().long()
        self.parents = parents.detach().cpu().long()
        self.voxels = voxels.detach().cpu()[:, :3].float()
        self.lod_offset, self.base_lod = list(lod_offset), base_lod
        self.dec = [tuple(t.detach().cpu().float() for t in d) for d in decoders]

    def features(self, x, lod, pidx):
        v = pidx.long() + self.lod_offset[lod]
        n = torch.addcmul(torch.full_like(x, 0.5), x, torch.full_like(x, 0.5))        # fmaf(x, .5, .5)
        chain = []
        for l in range(lod, -1, -1):
            chain.append((l, v))
            if l > 0:
                v = self.parents[v]
        feat = None
        for l, v in reversed(chain):                                                   # coarse -> fine
            res = float(1 << (l + self.base_lod))
            f = n * res - self.voxels[v]
            g = 1.0 - f
            s = 0
            for k in range(8):
                w = (f[:, 0] if k & 1 else g[:, 0]) * (f[:, 1] if k & 2 else g[:, 1]) * (f[:, 2] if k & 4 else g[:, 2])
                s = s + w.unsqueeze(1) * self.cf[self.trinkets[v, k]]
            feat = s if feat is None else s + feat
        return feat

    def sdf(self, x, lod, pidx):
        w0, b0, w1, b1 = self.dec[lod]
        feat = self.features(x, lod, pidx)
        inp = feat if self.pos_invariant else torch.cat([x, feat], dim=-1)
        return F.linear(F.relu(F.linear(inp, w0, b0)), w1, b1)


def spc_sphere_trace(snet, lod, nuggets, level_points, ray_o, ray_d, num_steps=50, min_dis=0.0003, far=5.0, h=0.001):
    """SDF::sphereTrace + getNormal, sol-renderer/SDF.cu:218-472, as the batch loop it is there.
    Returns dict(x, depth, hit, normal, pidx)."""
    level = lod + snet.base_lod
    n = ray_o.shape[0]
    x, t, cond, pidx = spc_ray_aabb(nuggets, level_points, level, ray_o, ray_d)        # :353-371 (init)
    has_run = torch.zeros(n, dtype=torch.bool)
    has_run[nuggets[:, 0].long().unique()] = True
    x = torch.where(has_run.unsqueeze(1), x, ray_o.float())
    d = torch.zeros(n, 1)
    dprev = torch.zeros(n, 1)
    hit = torch.zeros(n, dtype=torch.bool)
    for _ in range(num_steps):                                                         # :378
        act = cond.nonzero()[:, 0]
        if act.numel() == 0:
            break
        _d = snet.sdf(x[act], lod, pidx[act])                                          # :394-413
        d[act] = _d                                                                     # step.cuh:53
        t[act] = t[act] + _d
        h_ = (_d.abs()[:, 0].double() < float(np.float32(min_dis))) | \
             (((_d + dprev[act]).abs()[:, 0].double() * 0.5) < float(np.float32(min_dis * 5.0)))
        hit[act] = h_
        cond[act] = (t[act, 0] < far) & ~h_
        x[act] = torch.addcmul(ray_o[act], ray_d[act], t[act])                        # fmaf in the kernel; 1 ulp apart
        dprev[act] = _d
        x, t, cond, pidx = spc_ray_aabb(nuggets, level_points, level, ray_o, ray_d, query=x, active=cond,
                                        x=x, t=t, cond=cond, pidx=pidx)                # :442-460
    normal = torch.zeros(n, 3)
    hi = hit.nonzero()[:, 0]
    if hi.numel():
        g = []
        for k in range(3):
            e = torch.zeros(3)
            e[k] = h
            g.append(snet.sdf(x[hi] + e, lod, pidx[hi]) - snet.sdf(x[hi] - e, lod, pidx[hi]))
        g = torch.cat(g, dim=1)
        normal[hi] = g / g.norm(dim=1, keepdim=True).clamp_min(1e-30)
    return dict(x=x, depth=t, hit=hit, normal=normal, pidx=pidx)


# ---------------------------------------------------------------------------------------------------------------------
# Philox-4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11) -- the counter-based
# generator of the mesh sampler kernel (csrc/sample_mesh.cu).  The reference draws its samples with torch's host RNG
# (sdf-net/lib/torchgp/sample_surface.py:47-50, sample_uniform.py:31), so sampler parity is distributional; this
# restatement pins the kernel's stream to the published algorithm (Random123 known-answer vectors, tests/).
def philox4x32_10(counter, key):
    c, k = [int(v) & 0xffffffff for v in counter], [int(v) & 0xffffffff for v in key]
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xffffffff, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xffffffff]
        k = [(k[0] + 0x9E3779B9) & 0xffffffff, (k[1] + 0xBB67AE85) & 0xffffffff]
    return c


def sample_uniform_philox(seed, index):
    """The 'rand' sample `index` of nglod_sample_mesh for Philox key `seed`: (word >> 8) * 2^-24 * 2 - 1 per coordinate."""
    r = philox4x32_10([index & 0xffffffff, index >> 32, 0, 0], [seed & 0xffffffff, seed >> 32])
    return [((w >> 8) / 16777216.0) * 2.0 - 1.0 for w in r[:3]]
