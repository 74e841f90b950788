"""Where Renderer._shade_images_pipelined spends its wall time: cProfile of 100 calls + host-enqueue share."""
import copy, cProfile, io, pstats, sys, time, torch, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from nglod_b200.lib.renderer import Renderer
from nglod_b200.lib.tracer import SphereTracer
dev = torch.device('cuda', 0)
net, args = bench.build_and_fit(dev, lambda m: None)
a2 = copy.copy(args); a2.render_res = [bench.W, bench.H]
r = Renderer(SphereTracer(a2), args=a2, device=dev)
f = lambda: r.shade_images(net, f=bench.CAM_FROM, t=bench.CAM_TO, fov=bench.FOV)
for _ in range(6): f()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50): f()
print("wall per call %.3f ms" % ((time.perf_counter() - t0) / 50 * 1e3))
pr = cProfile.Profile(); pr.enable()
for _ in range(100): f()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue()[:5000])
