"""How long is one tile round when the machine is (nearly) idle?  Trace only the straggler rays of the bench frame."""
import sys, torch
sys.path.insert(0, '/root/repo')
import bench
from nglod_b200 import ops
from nglod_b200.lib.tracer import SphereTracer
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
tracer = SphereTracer(args)
def timed(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it * 1e3
print("whole frame: %.0f us" % timed(lambda: tracer(net, ray_o, ray_d)))
order = torch.argsort(steps, descending=True)
for cnt in (1, 15, 128, 512, 1056, 8843, 148 * 512):
    sel = order[:cnt]
    o, dd = ray_o[sel].contiguous(), ray_d[sel].contiguous()
    us = timed(lambda: tracer(net, o, dd))
    smax = int(steps[sel].max()); smin = int(steps[sel].min())
    print(f"{cnt:6d} longest rays (steps {smin}..{smax}): {us:8.1f} us  -> {us / (smax + 7):.2f} us per round of the longest ray")
# the same ray replicated: 1 .. 128 copies in one tile, and one copy per group over the whole machine
one = order[:1]
for copies in (1, 32, 128, 512, 148 * 512):
    o, dd = ray_o[one].repeat(copies, 1).contiguous(), ray_d[one].repeat(copies, 1).contiguous()
    us = timed(lambda: tracer(net, o, dd))
    print(f"longest ray x {copies:6d}: {us:8.1f} us -> {us / (int(steps[one]) + 7):.2f} us per round")
