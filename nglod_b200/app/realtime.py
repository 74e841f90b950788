"""Headless real-time loop -- the reference renderer's `display()` cycle without the GL window
(sol-renderer/sdfRenderer.cpp:217-242: update the camera, generate rays, traverse the octree, sphere-trace inside the
voxels, shade; camera set-up spc_raytrace_cuda.cpp:65-121).  Every stage of a frame is a device kernel of this
package and the frame stays on the device (an RGB buffer a presenter would blit):

    rays      nglod_generate_rays                     (camera orbiting the object)
    geometry  sparse: nglod_spc_raytrace_{count,fill} + nglod_spc_sphere_trace over a SparseOctreeSDF   (--spc-level L)
              dense : nglod_sphere_trace over the OctreeSDF                                             (default)
    shading   nglod_shade_matcap

    python -m nglod_b200.app.realtime --net OctreeSDF --num-lods 5 --pretrained m.pth --render-res 1920 1080 \
           --lod 4 [--spc-level 6] [--frames 120]

Without --pretrained a procedural torus is fitted first (there are no assets in the repo).  Prints per-frame device
time and fps; `run()` returns them so tests / bench.py can call it.
"""
import math
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

from .. import ops
from ..lib.options import parse_options
from ..lib.geoutils import _window, procedural_matcap
from ..lib.models import OctreeSDF  # noqa: F401  (resolved by name like the reference)
from ..lib.models import *  # noqa: F401,F403


def camera_basis(eye, target):
    """view / right / up of look_at (geoutils.py:180-188), on the host."""
    from ..lib.geoutils import camera_basis as _basis
    return _basis(eye, target)


def run(net, width, height, frames=60, lod=None, spc_level=None, fov=30.0, radius=None, height_y=2.8, device=None,
        num_steps=None, log=None):
    """Render `frames` frames of an orbit around the object.  Returns dict(ms=[per-frame device ms], fps, rgb, hit)
    with the last frame's device buffers."""
    device = device or next(net.parameters()).device
    lod = net.num_lods - 1 if lod is None else lod
    net.lod = lod
    radius = radius if radius is not None else math.hypot(2.8, 2.8)
    tan = np.float32(np.tan(np.radians(fov / 2)))
    with torch.cuda.device(device):
        wx, wy = _window(width, height, device)           # jittered window coordinates, drawn once (a real loop re-draws)
    matcap = procedural_matcap(device=device).tex
    sparse = None
    if spc_level is not None:
        from ..lib import spc as S
        from ..lib.torchgp import torus, normalize
        mesh = getattr(net, "_rt_mesh", None) or torus(0.6, 0.25, 128, 64)
        V, Fc = normalize(*[t.to(device) for t in mesh])
        sparse = S.SparseOctreeSDF(net, S.SPC(S.mesh_to_octree(V, Fc, spc_level, num_samples=1 << 22)))
        if lod + sparse.base_lod > spc_level:
            raise ValueError("--lod + base_lod must not exceed --spc-level")
    view_dev = net.net_view()
    ev = []
    rgb = hit = None
    for k in range(frames):
        a = 1.25 * math.pi + 2.0 * math.pi * k / max(frames, 1)     # frame 0 = the reference's default camera (-2.8, 2.8, -2.8)
        eye = [radius * math.cos(a), height_y, radius * math.sin(a)]
        origin, view, right, up = camera_basis(eye, [0.0, 0.0, 0.0])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ray_o, ray_d = ops.generate_rays(origin, view, right, up, tan, False, wx, wy)
        if sparse is not None:
            x, depth, hit, normal, _ = sparse.trace(ray_o, ray_d, lod)
        else:
            kw = {} if num_steps is None else {"num_steps": num_steps}
            x, depth, hit, normal = ops.sphere_trace(view_dev, lod, ray_o, ray_d, **kw)
        rgb = ops.shade_matcap(ray_d, normal, hit, matcap)
        e1.record()
        ev.append((e0, e1))
    torch.cuda.synchronize(device)
    ms = [a.elapsed_time(b) for a, b in ev]
    steady = ms[min(3, len(ms) - 1):] or ms
    out = {"ms": ms, "fps": 1e3 / float(np.mean(steady)), "rgb": rgb, "hit": hit,
           "mode": f"sparse level {spc_level}" if sparse is not None else "dense"}
    if log:
        log(f"[realtime] {width}x{height} lod {lod} {out['mode']}: {np.mean(steady):.3f} ms/frame = {out['fps']:.0f} fps "
            f"(min {min(ms):.3f}, max {max(ms):.3f} ms over {len(ms)} frames)")
    return out


def main(argv=None):
    parser = parse_options(return_parser=True)
    app = parser.add_argument_group("app")
    app.add_argument("--frames", type=int, default=120)
    app.add_argument("--spc-level", type=int, default=None, help="trace inside the voxels of an octree of this level")
    args = parser.parse_args(argv)
    if not torch.cuda.is_available():
        raise RuntimeError("nglod_b200 renders on a CUDA device only (no CPU path)")
    device = torch.device("cuda")
    net = globals()[args.net](args)
    if args.pretrained is not None:
        net.load_state_dict(torch.load(args.pretrained, map_location="cpu"))
        net.to(device).eval()
    else:
        from ..lib.datasets import MeshDataset
        from ..lib.trainer import FusedTrainer
        from ..lib.torchgp import torus
        net.to(device)
        ds = MeshDataset(args, mesh=torus(0.6, 0.25, 128, 64), device=device)
        tr = FusedTrainer(net, lr=1e-3)
        g = torch.Generator(device=device).manual_seed(7)
        for _ in range(300):
            idx = torch.randint(0, len(ds), (65536,), device=device, generator=g)
            tr.step(ds.pts[idx], ds.d[idx])
        net.eval()
    w, h = args.render_res
    t0 = time.time()
    out = run(net, w, h, frames=args.frames, lod=args.lod, spc_level=args.spc_level, fov=args.camera_fov, log=print)
    print(f"[realtime] {args.frames} frames in {time.time() - t0:.2f} s wall")
    return out


if __name__ == "__main__":
    main()
    sys.exit(0)
