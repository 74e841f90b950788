"""One line: the camera-driven host frame (trace_lookat_host(packed=True)), wall per frame back to back and with an L2 flush +
idle GPU before each frame; the same call's device time (events)."""
import sys, time, torch, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from nglod_b200.lib.tracer import SphereTracer
from nglod_b200.lib.geoutils import _window
dev = torch.device('cuda', 0)
net, args = bench.build_and_fit(dev, lambda m: None)
tracer = SphereTracer(args)
W, H = bench.W, bench.H
torch.manual_seed(1000)
wx, wy = _window(W, H, "cpu"); wx, wy = wx.pin_memory(), wy.pin_memory()
cam = bench.camera_from(0.0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = []
for fields in (("depth", "hit", "normal"), ("x", "depth", "hit", "normal")):
    out = {}
    def c(): return tracer.trace_lookat_host(net, cam, bench.CAM_TO, W, H, fov=bench.FOV, window=(wx, wy), out=out, fields=fields, packed=True)
    for _ in range(5): c()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(50): c()
    b2b = (time.perf_counter() - t0) / 50 * 1e3
    ts = []
    for _ in range(20):
        flush.zero_(); torch.cuda.synchronize()
        t0 = time.perf_counter(); c(); ts.append((time.perf_counter() - t0) * 1e3)
    res.append("%s: back-to-back %.3f ms, flushed %.3f ms" % ("+".join(fields), b2b, float(np.median(ts))))
print(" | ".join(res))
