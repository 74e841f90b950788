"""Histogram of march steps per ray for the bench frame (which rays set the frame's critical path?)."""
import sys, torch
sys.path.insert(0, '/root/repo')
import bench
from nglod_b200 import ops
dev = torch.device('cuda', 0)
net, args = bench.build_and_fit(dev, print)
ray_o, ray_d = bench.make_rays(dev)
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
        d[live] = net(x[live])
        t = torch.where(col, t + d, t)
        steps += live.int()
s = steps.cpu()
print("rays", s.numel(), "marched", int((s > 0).sum()), "mean steps (marched)", float(s[s > 0].float().mean()))
for thr in (8, 16, 32, 64, 128, 200, 255):
    print(f"  rays with > {thr:3d} steps: {int((s > thr).sum())}")
print("max", int(s.max()))
