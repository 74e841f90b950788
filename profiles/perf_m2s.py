import sys, time, torch
sys.path.insert(0, '/root/repo')
from nglod_b200 import ops
from nglod_b200.lib.torchgp import torus, point_sample, normalize
dev = 'cuda'
V, F = normalize(*[t.to(dev) for t in torus(0.6, 0.25, 128, 64)])
tri = V[F].contiguous()
def t(fn, it=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it
modes = ["rand", "near", "near", "trace", "trace"]
print("point_sample ms", t(lambda: point_sample(V, F, modes, 100000)))
pts = point_sample(V, F, modes, 100000)
print("mesh2sdf 500k x", tri.shape[0], "ms", t(lambda: ops.mesh2sdf_gpu(pts, tri)))
for name, sl in (("rand", slice(0, 100000)), ("near", slice(100000, 300000)), ("trace", slice(300000, 500000))):
    p = pts[sl].contiguous()
    print(" ", name, p.shape[0], "ms", t(lambda: ops.mesh2sdf_gpu(p, tri)))
sys.path.insert(0, '/root/repo/oracle')
import build_ref
ref = build_ref.load_ref("ref_mesh2sdf")
if ref is not None:
    print("reference mesh2sdf_gpu (its own kernels, same box) 500k ms", t(lambda: ref.mesh2sdf_gpu(pts, tri), it=2))
    refa = build_ref.load_ref("ref_sol_nglod")
    from nglod_b200.lib.geoutils import look_at
    ro, rd = look_at([-2.8, 2.8, -2.8], [0, 0, 0], 1280, 720, mode="persp", fov=30.0, device=dev)
    print("reference aabb 720p ms", t(lambda: refa.aabb(ro, rd), it=20), " ours ms", t(lambda: ops.aabb(ro, rd), it=20))
