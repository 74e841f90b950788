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
cfgs = [dict(chunks=c, streams=s) for c in (3, 4, 5, 6) for s in (2, 3)] + [dict(fractions=f, streams=s) for f in ((0.35, 0.65, 0.85), (0.3, 0.55, 0.75, 0.9), (0.4, 0.7)) for s in (2, 3)]
for cfg in cfgs:
    ch = cfg
    for _ in range(3): rb = tr.trace_host(net, ho, hd, out=out, **cfg)
    assert torch.equal(rb.depth, ref.depth.cpu()) and torch.equal(rb.hit, ref.hit.cpu())
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): tr.trace_host(net, ho, hd, out=out, **cfg)
    dt = (time.perf_counter() - t0) / 20
    print("chunks", ch, f"{dt*1e3:.3f} ms  {n/dt:.3e} rays/s")
