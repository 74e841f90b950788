import sys, time, torch
sys.path.insert(0, "/root/repo")
import bench
from nglod_b200 import ops
from nglod_b200.lib.tracer import SphereTracer
dev = torch.device("cuda", 0)
net, args = bench.build_and_fit(dev, print)
o, d = bench.make_rays(dev)
ho, hd = o.cpu().pin_memory(), d.cpu().pin_memory()
tr = SphereTracer(args)
n = o.shape[0]
out = {"x": torch.empty(n, 3).pin_memory(), "depth": torch.empty(n, 1).pin_memory(), "hit": torch.empty(n, dtype=torch.bool).pin_memory(), "normal": torch.empty(n, 3).pin_memory()}
def T(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(it): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / it * 1e3
od, dd = torch.empty_like(o), torch.empty_like(d)
print("h2d only", T(lambda: (od.copy_(ho, non_blocking=True), dd.copy_(hd, non_blocking=True))))
rb = tr(net, o, d)
print("d2h only", T(lambda: [out[k].copy_(getattr(rb, k), non_blocking=True) for k in out]))
print("kernel only", T(lambda: tr(net, o, d)))
view = net.net_view()
def chunked(ch):
    b = [(n * i) // ch for i in range(ch + 1)]
    for i in range(ch): ops.sphere_trace(view, 4, o[b[i]:b[i+1]], d[b[i]:b[i+1]])
for ch in (2, 4, 8): print("kernel in", ch, "chunks", T(lambda: chunked(ch)))
s2 = torch.cuda.Stream(dev)
def both():
    with torch.cuda.stream(s2):
        od.copy_(ho, non_blocking=True); dd.copy_(hd, non_blocking=True)
    tr(net, o, d)
print("kernel + concurrent h2d on another stream", T(both))
def both2():
    with torch.cuda.stream(s2):
        [out[k].copy_(getattr(rb, k), non_blocking=True) for k in out]
    tr(net, o, d)
print("kernel + concurrent d2h on another stream", T(both2))
