"""[historical: the NGLOD_TRACE_HOLD knob existed only for this sweep; the result is recorded in profiles/README.md]
Straggler hold (NGLOD_TRACE_HOLD = march steps after which a group stops refilling): frame time and identical output."""
import os, sys, torch
sys.path.insert(0, '/root/repo')
import bench
from nglod_b200.lib.tracer import SphereTracer
dev = torch.device('cuda', 0)
net, args = bench.build_and_fit(dev, print)
ray_o, ray_d = bench.make_rays(dev)
tracer = SphereTracer(args)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(it):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / it * 1e3
ref = None
for lod in (4,):
    net.lod = lod
    ref = None
    for hold in (0, 64, 80, 96, 112, 128, 160, 192, 0):
        os.environ["NGLOD_TRACE_HOLD"] = str(hold)
        rb = tracer(net, ray_o, ray_d)
        out = (rb.x.clone(), rb.depth.clone(), rb.hit.clone(), rb.normal.clone())
        if ref is None: ref = out
        same = all(torch.equal(a, b) for a, b in zip(ref, out))
        print(f"lod {lod} hold {hold:3d}: {timed(lambda: tracer(net, ray_o, ray_d)):7.1f} us  identical {same}")
