import sys, time, torch
sys.path.insert(0, "/root/repo")
import bench
from nglod_b200.lib.tracer import SphereTracer
dev = torch.device("cuda", 0)
net, args = bench.build_and_fit(dev, print)
o, d = bench.make_rays(dev)
ho, hd = o.cpu().pin_memory(), d.cpu().pin_memory()
tr = SphereTracer(args)
n = o.shape[0]
out = {"x": torch.empty(n, 3).pin_memory(), "depth": torch.empty(n, 1).pin_memory(), "hit": torch.empty(n, dtype=torch.bool).pin_memory(), "normal": torch.empty(n, 3).pin_memory()}
ref = tr(net, o, d)
import itertools
cfgs = [dict(chunks=3, streams=2)] + [dict(fractions=f, streams=2) for f in ((0.1, 0.45, 0.8), (0.15, 0.5, 0.8), (0.2, 0.6), (0.08, 0.3, 0.6, 0.85), (0.25, 0.6, 0.9))]
for cfg in cfgs:
    ch = cfg
    for _ in range(3): rb = tr.trace_host(net, ho, hd, out=out, **cfg)
    assert torch.equal(rb.depth, ref.depth.cpu()) and torch.equal(rb.hit, ref.hit.cpu())
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): tr.trace_host(net, ho, hd, out=out, **cfg)
    dt = (time.perf_counter() - t0) / 20
    print("chunks", ch, f"{dt*1e3:.3f} ms  {n/dt:.3e} rays/s")
