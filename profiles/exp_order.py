"""Does the ray order matter? (the persistent tracer hands rays out in array order; the frame ends with its stragglers)"""
import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from nglod_b200 import ops
from nglod_b200.lib.tracer import SphereTracer
dev = torch.device('cuda', 0)
net, args = bench.build_and_fit(dev, print)
ray_o, ray_d = bench.make_rays(dev)
tracer = SphereTracer(args)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
# per-ray march steps (generic loop)
x, t, live = ops.aabb(ray_o, ray_d)
steps = torch.zeros(ray_o.shape[0], dtype=torch.int32, device=dev)
with torch.no_grad():
    d = net(x); dprev = d.clone()
    for i in range(256):
        flag = (t.abs() < 10.0)[:, 0]
        live = live & (d.abs() > 3e-4)[:, 0] & (((d + dprev) / 2).abs() > 9e-4)[:, 0] & flag
        if not bool(live.any()): break
        col = live.unsqueeze(1)
        x = torch.where(col, torch.addcmul(ray_o, ray_d, t), x)
        dprev = torch.where(col, d, dprev)
        d[live] = net(x[live]); t = torch.where(col, t + d, t); steps += live.int()
def timeit(o, dd):
    for _ in range(3): tracer(net, o, dd)
    ts = []
    for _ in range(15):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); tracer(net, o, dd); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.mean(ts))
g = torch.Generator(device=dev).manual_seed(0)
orders = {"image order (x-major)": torch.arange(ray_o.shape[0], device=dev),
          "reversed": torch.arange(ray_o.shape[0] - 1, -1, -1, device=dev),
          "random permutation": torch.randperm(ray_o.shape[0], device=dev, generator=g),
          "longest rays first (oracle order)": torch.argsort(steps, descending=True),
          "longest rays last": torch.argsort(steps, descending=False)}
for name, p in orders.items():
    print(f"{name:36s} {timeit(ray_o[p].contiguous(), ray_d[p].contiguous()):.4f} ms")
