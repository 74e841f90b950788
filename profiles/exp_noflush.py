"""Tracer frame time with and without the L2 flush between iterations (how much do cold ray / grid loads cost?)."""
import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from nglod_b200.lib.tracer import SphereTracer
dev = torch.device('cuda', 0)
net, args = bench.build_and_fit(dev, print)
ray_o, ray_d = bench.make_rays(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tracer = SphereTracer(args)
for do_flush in (True, False):
    for _ in range(3): tracer(net, ray_o, ray_d)
    ts = []
    for _ in range(20):
        if do_flush: flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); tracer(net, ray_o, ray_d); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f"flush={do_flush}: tracer {np.mean(ts):.4f} ms (min {min(ts):.4f})  {ray_o.shape[0]/np.mean(ts)*1e3:.3e} rays/s")
