"""Generate golden vectors by running the UNMODIFIED reference (read-only at /root/reference).

Run in the authoring container only:   python tests/golden/make_golden.py
The reference ships no tests or fixtures, so these files are how parity is
pinned: the reference's own classes (OctreeSDF, SphereTracer, Renderer, look_at, gradient) are
imported in place, with empty stub modules for third-party packages that are not installed and are
not on the path (polyscope, tinyobjloader, pyexr, moviepy, matplotlib), `lib.utils.PerfTimer`
replaced (its constructor needs a CUDA driver), and `sol_nglod.aabb` -- which exists only as CUDA --
provided by oracle/oracle.c (itself checked against the compiled reference kernel on the GPU box).

Outputs (small .npz files next to this script):
  rand5.npz  random-init 5-LOD model re-creatable from the seed: sdf at every LOD, return_lst,
             gradients of the L2 loss, a 64x36 trace at lod 4, a mixed-origin ray set
  fit3.npz   a 3-LOD model fitted for a few hundred Adam steps to a torus (weights stored):
             sdf, 96x54 trace at lod 2, Renderer.render with AO
  samplers.npz  lib/torchgp's host sampling recipe under fixed torch seeds
  fit5.npz   BASELINE config 2's model shape: 5 LODs (base-lod 2: R = 4..64), feature-dim 32, fitted to a torus, evaluated
             at lod 4: sdf, a 96x54 SphereTracer frame, SphereTracer.get_min, Renderer.render with --shadow
             --ground-height -0.3 --ao (shadow-ray jitter stored).  To keep the fixture small the two fine grids start
             at ZERO and are trained only on near-surface samples in a second phase (Adam leaves untouched nodes at
             exactly 0, which compresses away), and every weight is rounded to an fp16-exact value before the reference
             is evaluated (stored as fp16: ~1.5 MB instead of 40 MB).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = "/root/reference/sdf-net"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import nglod_oracle as O  # noqa: E402

for m in ["polyscope", "tinyobjloader", "pyexr", "moviepy", "moviepy.editor", "matplotlib", "matplotlib.pyplot",
          "sol_nglod", "mesh2sdf", "cv2"]:
    sys.modules.setdefault(m, types.ModuleType(m))
sys.modules["sol_nglod"].aabb = O.aabb

import lib.utils  # noqa: E402


class _NoTimer:
    def __init__(self, activate=False):
        pass

    def check(self, name=None):
        pass

    def reset(self):
        pass


lib.utils.PerfTimer = _NoTimer
from lib.options import parse_options  # noqa: E402
from lib.models import OctreeSDF  # noqa: E402
from lib.tracer import SphereTracer  # noqa: E402
from lib.renderer import Renderer  # noqa: E402
from lib.geoutils import look_at  # noqa: E402
from lib.diffutils import gradient  # noqa: E402

torch.set_num_threads(8)


def make_args(extra):
    return parse_options(return_parser=True).parse_args(["--net", "OctreeSDF", "--feature-dim", "32"] + extra)


def torus_sdf(p, major=0.6, minor=0.25):
    q = torch.sqrt(p[..., 0] ** 2 + p[..., 2] ** 2) - major
    return torch.sqrt(q * q + p[..., 1] ** 2) - minor


def trace_pack(prefix, rb, out):
    out[prefix + "_x"] = rb.x.detach().numpy()
    out[prefix + "_depth"] = rb.depth.detach().numpy()
    out[prefix + "_hit"] = rb.hit.detach().numpy()
    out[prefix + "_normal"] = rb.normal.detach().numpy()


def rand5():
    out = {}
    args = make_args(["--num-lods", "5"])
    torch.manual_seed(0)
    net = OctreeSDF(args)
    out["weights_checksum"] = np.array([float(sum(p.double().sum() for p in net.parameters())),
                                        float(sum((p.double() ** 2).sum() for p in net.parameters()))])
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(4096, 3, generator=g) * 2 - 1) * 1.2          # some points outside the unit box
    x[:8] = torch.tensor([[1.0, 1.0, 1.0], [-1.0, -1.0, -1.0], [0.0, 0.0, 0.0], [1.0, 0.0, -1.0],
                          [0.5, 0.5, 0.5], [-0.25, 0.75, 1.0], [2.0, -2.0, 0.3], [1.0 - 1e-7, 0.0, 0.0]])
    out["x"] = x.numpy()
    net.eval()
    with torch.no_grad():
        for l in range(5):
            out[f"sdf_lod{l}"] = net.sdf(x, lod=l).numpy()
        lst = net.sdf(x, return_lst=True)
        out["sdf_lst"] = np.stack([t.numpy() for t in lst])
        net.lod = 3
        out["forward_lod3"] = net(x).numpy()
        net.lod = None
    # gradients of the reference training objective at lod 4 (trainer.py:317-339) and of a 2-LOD sum
    gt = torus_sdf(x).unsqueeze(1)
    out["gt"] = gt.numpy()
    rng = np.random.RandomState(7)
    for tag, lods in (("g4", [4]), ("g13", [1, 3])):
        net.zero_grad()
        loss = 0
        for l in lods:
            loss = loss + ((net.sdf(x, lod=l) - gt) ** 2).sum()
        loss = loss / x.shape[0]
        loss.backward()
        out[f"{tag}_loss"] = np.array(loss.item())
        for i in range(5):
            gfm = net.features[i].fm.grad
            if gfm is None:
                continue
            gl = gfm[0].permute(1, 2, 3, 0).contiguous().reshape(-1)     # channels-last flat, like the kernels
            if i <= 1:
                out[f"{tag}_fm{i}"] = gl.numpy()
            else:
                idx = rng.randint(0, gl.numel(), size=4096)
                out[f"{tag}_fm{i}_idx"] = idx
                out[f"{tag}_fm{i}_val"] = gl.numpy()[idx]
            out[f"{tag}_fm{i}_sum"] = np.array([gl.double().sum().item(), gl.double().abs().sum().item()])
        for l in lods:
            for k in ("0.weight", "0.bias", "2.weight", "2.bias"):
                out[f"{tag}_louts{l}.{k}"] = dict(net.louts[l].named_parameters())[k].grad.numpy()
        out[f"{tag}_untouched_head_has_grad"] = np.array(
            [net.louts[0][0].weight.grad is not None and float(net.louts[0][0].weight.grad.abs().sum()) > 0])
    # autodiff d sdf / d x (diffutils.py:32-37) incl. the border-clip rule
    xg = x[:512].clone()
    net.lod = 4
    out["autodiff_lod4"] = gradient(xg, net, method="autodiff").detach().numpy()
    out["finitediff_lod4"] = gradient(x[:512].clone(), net, method="finitediff").detach().numpy()
    # tracer, lod 4, default options, seeded jitter
    tracer = SphereTracer(args)
    torch.manual_seed(123)
    ray_o, ray_d = look_at([-2.8, 2.8, -2.8], [0, 0, 0], 64, 36, fov=30.0, mode="persp", device="cpu")
    out["t1_ray_o"], out["t1_ray_d"] = ray_o.numpy(), ray_d.numpy()
    trace_pack("t1", tracer(net, ray_o, ray_d), out)
    # mixed origins (inside / outside the box, quirk 6), short budget
    g2 = torch.Generator().manual_seed(5)
    ro = torch.rand(512, 3, generator=g2) * 3 - 1.5
    rd = torch.nn.functional.normalize(torch.randn(512, 3, generator=g2), dim=1)
    tr12 = SphereTracer(args, num_steps=12)
    out["t2_ray_o"], out["t2_ray_d"] = ro.numpy(), rd.numpy()
    trace_pack("t2", tr12(net, ro, rd), out)
    np.savez_compressed(os.path.join(HERE, "rand5.npz"), **out)
    print("rand5:", {k: v.shape for k, v in out.items() if k.startswith("t1_")}, "hits", out["t1_hit"].sum(), out["t2_hit"].sum())


def fit3():
    out = {}
    args = make_args(["--num-lods", "3"])
    torch.manual_seed(3)
    net = OctreeSDF(args)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    net.train()
    g = torch.Generator().manual_seed(11)
    for it in range(400):
        pts = torch.rand(8192, 3, generator=g) * 2 - 1
        gt = torus_sdf(pts).unsqueeze(1)
        opt.zero_grad()
        preds = net.sdf(pts, return_lst=True)
        loss = sum(((p - gt) ** 2).sum() for p in preds) / pts.shape[0]
        loss.backward()
        opt.step()
        if it % 100 == 0:
            print("fit3 it", it, loss.item())
    net.eval()
    for k, v in net.state_dict().items():
        out["sd." + k] = v.numpy()
    x = torch.rand(4096, 3, generator=g) * 2 - 1
    out["x"] = x.numpy()
    with torch.no_grad():
        for l in range(3):
            out[f"sdf_lod{l}"] = net.sdf(x, lod=l).numpy()
    net.lod = 2
    tracer = SphereTracer(args)
    torch.manual_seed(321)
    ray_o, ray_d = look_at([-2.8, 2.8, -2.8], [0, 0, 0], 96, 54, fov=30.0, mode="persp", device="cpu")
    out["t1_ray_o"], out["t1_ray_d"] = ray_o.numpy(), ray_d.numpy()
    rb = tracer(net, ray_o, ray_d)
    trace_pack("t1", rb, out)
    with torch.no_grad():
        conv = (net(rb.x).abs() < 0.0003)[:, 0] & rb.hit
    out["t1_converged"] = conv.numpy()
    # Renderer.render with ambient occlusion (renderer.py:109-216), identical rays
    rargs = make_args(["--num-lods", "3", "--render-res", "96", "54", "--ao"])
    renderer = Renderer(SphereTracer(rargs), args=rargs, device="cpu")
    rb2 = renderer.render(net, ray_o, ray_d)
    out["r1_ao"] = rb2.ao.detach().numpy()
    out["r1_relative_depth"] = rb2.relative_depth.detach().numpy()
    out["r1_hit"] = rb2.hit.detach().numpy()
    out["r1_normal"] = rb2.normal.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "fit3.npz"), **out)
    print("fit3: hits", int(rb.hit.sum()), "converged", int(conv.sum()), "of", rb.hit.numel())




def fit5():
    out = {}
    args = make_args(["--num-lods", "5"])
    torch.manual_seed(5)
    net = OctreeSDF(args)
    with torch.no_grad():
        net.features[3].fm.zero_()
        net.features[4].fm.zero_()
    g = torch.Generator().manual_seed(13)

    def near_surface(nb):
        """points within a few hundredths of the torus surface (the 'near' / 'trace' sample modes of MeshDataset)"""
        u = torch.rand(nb, generator=g) * 2 * np.pi
        v = torch.rand(nb, generator=g) * 2 * np.pi
        ring = torch.stack([0.6 * torch.cos(u), torch.zeros(nb), 0.6 * torch.sin(u)], 1)
        nrm = torch.stack([torch.cos(v) * torch.cos(u), torch.sin(v), torch.cos(v) * torch.sin(u)], 1)
        return ring + nrm * (0.25 + 0.02 * torch.randn(nb, 1, generator=g))

    net.train()
    # phase 1: the three coarse grids + every decoder, uniform + near-surface samples
    p1 = [net.features[i].fm for i in range(3)] + list(net.louts.parameters())
    opt = torch.optim.Adam(p1, lr=1e-3)
    for it in range(250):
        pts = torch.cat([torch.rand(4096, 3, generator=g) * 2 - 1, near_surface(4096)])
        gt = torus_sdf(pts).unsqueeze(1)
        opt.zero_grad()
        preds = net.sdf(pts, return_lst=True)
        loss = sum(((p - gt) ** 2).sum() for p in preds) / pts.shape[0]
        loss.backward()
        net.features[3].fm.grad = None
        net.features[4].fm.grad = None
        opt.step()
        if it % 50 == 0:
            print("fit5 phase 1 it", it, loss.item(), flush=True)
    # phase 2: everything, near-surface samples only -> the fine grids change in a thin band, the rest stays 0
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    for it in range(250):
        pts = near_surface(8192)
        gt = torus_sdf(pts).unsqueeze(1)
        opt.zero_grad()
        preds = net.sdf(pts, return_lst=True)
        loss = sum(((p - gt) ** 2).sum() for p in preds) / pts.shape[0]
        loss.backward()
        opt.step()
        if it % 50 == 0:
            print("fit5 phase 2 it", it, loss.item(), flush=True)
    net.eval()
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(p.half().float())                           # fp16-exact weights: what is stored is what was evaluated
    for k, v in net.state_dict().items():
        out["sd." + k] = v.half().numpy()
    print("fit5: non-zero fraction of the fine grids",
          [float((net.features[i].fm != 0).float().mean()) for i in (3, 4)], flush=True)
    x = torch.cat([torch.rand(2048, 3, generator=g) * 2 - 1, near_surface(2048)])
    out["x"] = x.numpy()
    with torch.no_grad():
        for l in range(5):
            out[f"sdf_lod{l}"] = net.sdf(x, lod=l).numpy()
    net.lod = 4
    tracer = SphereTracer(args)
    torch.manual_seed(654)
    ray_o, ray_d = look_at([-2.8, 2.8, -2.8], [0, 0, 0], 96, 54, fov=30.0, mode="persp", device="cpu")
    out["t1_ray_o"], out["t1_ray_d"] = ray_o.numpy(), ray_d.numpy()
    rb = tracer(net, ray_o, ray_d)
    trace_pack("t1", rb, out)
    with torch.no_grad():
        conv = (net(rb.x).abs() < 0.0003)[:, 0] & rb.hit
    out["t1_converged"] = conv.numpy()
    print("fit5: hits", int(rb.hit.sum()), "converged", int(conv.sum()), "of", rb.hit.numel(), flush=True)
    # SphereTracer.get_min (SphereTracer.py:134-218).  As shipped it ends in RenderBuffer(minx=...), a keyword the buffer
    # does not have (its field is min_x) -> TypeError; the golden run swaps in a buffer class that accepts it.
    ST = sys.modules["lib.tracer.SphereTracer"]      # the module (lib.tracer re-exports the class under the same name)

    class _AnyBuffer:
        def __init__(self, **kw):
            self.__dict__.update(kw)
    keep = ST.RenderBuffer
    ST.RenderBuffer = _AnyBuffer
    try:
        gm = SphereTracer(args, num_steps=48).get_min(net, ray_o, ray_d)
    finally:
        ST.RenderBuffer = keep
    for k in ("x", "depth", "hit", "normal", "minx"):
        out["gm_" + k] = getattr(gm, k).detach().numpy()
    # Renderer.render with the shadow pass and AO (renderer.py:131-209); the only random draw inside is the shadow-ray
    # jitter normal_(0, 0.01), replayed here from the same seed and stored
    rargs = make_args(["--num-lods", "5", "--render-res", "96", "54", "--shadow", "--ground-height", "-0.3", "--ao"])
    renderer = Renderer(SphereTracer(rargs), args=rargs, device="cpu")
    torch.manual_seed(777)
    rb2 = renderer.render(net, ray_o, ray_d)
    torch.manual_seed(777)
    out["r2_shadow_jitter"] = torch.zeros(ray_o.shape[0], 3).normal_(0.0, 0.01).numpy()
    for k in ("x", "depth", "hit", "normal", "shadow", "ao", "relative_depth"):
        out["r2_" + k] = getattr(rb2, k).detach().numpy()
    np.savez_compressed(os.path.join(HERE, "fit5.npz"), **out)
    print("fit5: shadow pixels", int(rb2.shadow.sum()), "plane+surface hits", int(rb2.hit.sum()),
          "file", os.path.getsize(os.path.join(HERE, "fit5.npz")) // 1024, "KB", flush=True)


def spc_golden():
    """Octree bytes / points / pyramid / point queries from the reference's numpy SPC (sdf-net/lib/spc3d.py), for a
    sphere of radius 0.6 at level 4 (recursive construction :128-137, binary encoding :224-255, decoding :273-304,
    Identify :309-338)."""
    from lib.spc3d import SPC3D
    level = 4
    spc = SPC3D(level)
    spc.construct(lambda x, y, z: np.sqrt(x * x + y * y + z * z) - 0.6)
    n = spc.psize
    leaf = spc.pdata[:n].copy()
    osize = spc.points_to_nodes()
    octree = np.array(spc.oroot[:osize], dtype=np.uint8).copy()
    spc2 = SPC3D(level, nodebytes=octree.copy())
    spc2.osize = octree.size
    spc2.oroot = spc2.odata
    num = spc2.nodes_to_points()
    pyramid = np.array(spc2.pyramid[:level + 1]).copy()
    total = int(pyramid.sum())
    mort = np.array(spc2.mdata[:total]).copy()
    rng = np.random.RandomState(3)
    q = rng.randint(-1, 17, size=(300, 3))
    ident = np.array([spc2.Identify(p) for p in q])
    np.savez_compressed(os.path.join(HERE, "spc.npz"), level=np.array(level), leaf_points=leaf, octree=octree,
                        pyramid=pyramid, morton_all=mort, query_pts=q, query_idx=ident)
    print("spc:", n, "leaf voxels,", osize, "octree bytes, pyramid", pyramid, "queries hit", int((ident >= 0).sum()))


def samplers():
    """The host sampling recipe of the reference (sdf-net/lib/torchgp: normalize, per_face_normals, point_sample,
    sample_surface, sample_near_surface, sample_spc) on a small procedural torus, under fixed torch seeds: the CPU path
    of nglod_b200.lib.torchgp must reproduce these bit for bit; the sampler KERNEL is compared with this recipe
    distributionally (its random stream is Philox, not torch's)."""
    import lib.torchgp as R
    from nglod_b200.lib.torchgp import torus
    V, F = torus(0.6, 0.25, 16, 8)
    V = V * 1.7 + 0.3                                           # not normalised yet
    Vn, Fn = R.normalize(V.clone(), F.clone())
    out = {"V": V.numpy(), "F": F.numpy(), "Vn": Vn.numpy(), "face_normals": R.per_face_normals(Vn, Fn).numpy()}
    torch.manual_seed(7)
    out["point_sample"] = R.point_sample(Vn, Fn, ["rand", "near", "trace", "near"], 64).numpy()
    torch.manual_seed(8)
    sp, sn = R.sample_surface(Vn, Fn, 50)
    out["surface_pts"], out["surface_nrm"] = sp.numpy(), sn.numpy()
    torch.manual_seed(9)
    out["near"] = R.sample_near_surface(Vn, Fn, 50, variance=0.02).numpy()
    corners = torch.tensor([[0, 0, 0], [7, 7, 7], [3, 1, 6], [2, 2, 2]], dtype=torch.int16)
    torch.manual_seed(10)
    out["spc_corners"], out["spc_samples"] = corners.numpy(), R.sample_spc(corners, 3, 6).numpy()
    np.savez_compressed(os.path.join(HERE, "samplers.npz"), **out)
    print("samplers:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    which = sys.argv[1:] or ["rand5", "fit3", "spc", "samplers", "fit5"]
    if "samplers" in which:
        samplers()
    if "rand5" in which:
        rand5()
    if "fit3" in which:
        fit3()
    if "spc" in which:
        spc_golden()
    if "fit5" in which:
        fit5()
