"""[NOT YET RUN: written at the end of round 1 without GPU minutes left]  Express-lane tracer experiment.
Build with   NGLOD_EXTRA_NVCC_FLAGS=-DNGLOD_TRACE_EXPRESS=1 python -c "from nglod_b200.build import build_library; build_library()"
then run this under `timeout 120` (a protocol bug would spin): sweeps NGLOD_TRACE_EXPRESS_CTAS x NGLOD_TRACE_HANDOFF on the
bench frame, checks that every output equals the run without the express lane (CTAS = 0), prints the frame time."""
import os, sys, torch
sys.path.insert(0, '/root/repo')
import bench
from nglod_b200.lib.tracer import SphereTracer
dev = torch.device('cuda', 0)
net, args = bench.build_and_fit(dev, print)
ray_o, ray_d = bench.make_rays(dev)
tracer = SphereTracer(args)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, it=10):
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
for ctas, hand in ((0, 32), (8, 32), (16, 32), (24, 32), (16, 16), (16, 48), (16, 64), (32, 24)):
    os.environ["NGLOD_TRACE_EXPRESS_CTAS"] = str(ctas); os.environ["NGLOD_TRACE_HANDOFF"] = str(hand)
    rb = tracer(net, ray_o, ray_d)
    out = (rb.x.clone(), rb.depth.clone(), rb.hit.clone(), rb.normal.clone())
    if ref is None: ref = out
    same = all(torch.equal(a, b) for a, b in zip(ref, out))
    print(f"express CTAs {ctas:2d} hand-off at step {hand:2d}: {timed(lambda: tracer(net, ray_o, ray_d)):7.1f} us  identical {same}", flush=True)
