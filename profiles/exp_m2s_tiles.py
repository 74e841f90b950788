"""mesh2sdf with / without the tile cull: bit-identical results, timings per sampling mode and batch size."""
import os, sys, torch
sys.path.insert(0, '/root/repo')
from nglod_b200 import ops
from nglod_b200.lib.torchgp import torus, icosphere, point_sample, normalize
dev = 'cuda'
def t(fn, it=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it
def both(pts, tri):
    d1 = ops.mesh2sdf_gpu(pts, tri)[0]; ms1 = t(lambda: ops.mesh2sdf_gpu(pts, tri))
    d0 = ops.mesh2sdf_gpu(pts, tri, force_walk=True)[0]; ms0 = t(lambda: ops.mesh2sdf_gpu(pts, tri, force_walk=True))
    return ms1, ms0, torch.equal(d1.view(torch.int32), d0.view(torch.int32)), int((d1 < 0).sum())
modes = ["rand", "near", "near", "trace", "trace"]
for name, (V, F) in (("torus128x64", torus(0.6, 0.25, 128, 64)), ("ico5", icosphere(5)), ("ico3", icosphere(3))):
    V, F = normalize(V.to(dev), F.to(dev))
    tri = V[F].contiguous()
    pts = point_sample(V, F, modes, 100000)
    print(name, "tris", tri.shape[0], "500k: cull %.2f ms, no cull %.2f ms, identical %s, inside %d" % both(pts, tri))
    for nm, sl in (("rand", slice(0, 100000)), ("near", slice(100000, 300000)), ("trace", slice(300000, 500000))):
        print("   ", nm, "cull %.2f ms, no cull %.2f ms, identical %s, inside %d" % both(pts[sl].contiguous(), tri))
    for n in (62500, 5000, 777):
        print("   ", n, "cull %.2f ms, no cull %.2f ms, identical %s, inside %d" % both(pts[torch.randperm(500000, device=dev)[:n]].contiguous(), tri))
# shuffled triangle order and a mesh with degenerate / duplicated triangles
V, F = normalize(*[x.to(dev) for x in torus(0.6, 0.25, 64, 32)])
tri = V[F].contiguous()
tri = torch.cat([tri, tri[:100], tri[:50, :1].expand(-1, 3, -1)], 0)[torch.randperm(tri.shape[0] + 150, device=dev)].contiguous()
pts = (torch.rand(200000, 3, device=dev) * 2.4 - 1.2)
print("degenerate+dup, points outside the box: cull %.2f ms, no cull %.2f ms, identical %s, inside %d" % both(pts, tri))
