"""A/B of the first-step value re-use in the tracer: frame time, evaluations per ray, and the frame itself (bitwise) against a
frame saved by the other build."""
import os, sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from nglod_b200 import ops
dev = torch.device('cuda', 0)
net, args = bench.build_and_fit(dev, lambda m: None)
ray_o, ray_d = bench.make_rays(dev)
view = net.net_view()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
stats = torch.zeros(2, dtype=torch.int64, device=dev)
x, depth, hit, normal = ops.sphere_trace(view, bench.LOD, ray_o, ray_d, stats=stats)
torch.cuda.synchronize()
ts = []
for _ in range(15):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ops.sphere_trace(view, bench.LOD, ray_o, ray_d); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
print("frame %.4f ms, evals/ray %.4f, hits %d, depth checksum %.6f" % (np.median(ts), stats[0].item() / ray_o.shape[0], int(hit.sum()), float(depth.double().sum())))
path = "/root/repo/gpurun_out/_reuse_frame.pt"
cur = {"x": x.cpu(), "depth": depth.cpu(), "hit": hit.cpu(), "normal": normal.cpu()}
if os.path.exists(path):
    ref = torch.load(path)
    for k in cur:
        same = torch.equal(cur[k], ref[k])
        diff = (cur[k].double() - ref[k].double()).abs()
        print(f"  {k}: identical {same}, differing entries {int((diff > 0).sum())}, max |diff| {float(diff.max()):.3e}")
else:
    torch.save(cur, path)
