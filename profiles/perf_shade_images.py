"""Renderer.shade_images at 1280x720 (the call SURVEY 8d quotes config 2's fps on): generic path (shade_tensor + cpu()) against
the pipelined one, by number of chunks.  Wall clock per call, steady state."""
import copy, sys, time, torch, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from nglod_b200.lib.renderer import Renderer
from nglod_b200.lib.tracer import SphereTracer
dev = torch.device('cuda', 0)
net, args = bench.build_and_fit(dev, lambda m: None)
a2 = copy.copy(args); a2.render_res = [bench.W, bench.H]
r = Renderer(SphereTracer(a2), args=a2, device=dev)
def wall(fn, it=20):
    for _ in range(6): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(it):
        t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts))
r.pipelined = False
print("generic: %.3f ms" % wall(lambda: r.shade_images(net, f=bench.CAM_FROM, t=bench.CAM_TO, fov=bench.FOV)))
r.pipelined = True
for c in (1, 2, 3, 4, 6, 8):
    r._pipe_ws = None
    print("pipelined, %d chunks: %.3f ms" % (c, wall(lambda: r._shade_images_pipelined(net, bench.CAM_FROM, bench.CAM_TO, bench.FOV, chunks=c))))
r.pipe_reserve_sms = 8
for fr in ((0.4, 0.8), (0.45, 0.85), (0.5, 0.9), (0.35, 0.65, 0.9), (0.4, 0.7, 0.92), (0.3, 0.6, 0.85, 0.95), (0.5, 0.8, 0.95), (0.6, 0.9)):
    r._pipe_ws = None
    print("reserve 8, split %s: %.3f ms" % (fr, wall(lambda: r._shade_images_pipelined(net, bench.CAM_FROM, bench.CAM_TO, bench.FOV, chunks=fr))))
