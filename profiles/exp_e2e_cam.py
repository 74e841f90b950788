"""Camera-driven host call (SphereTracer.trace_lookat_host): ranges x streams sweep on the bench frame."""
import sys, time, torch
sys.path.insert(0, '/root/repo')
import bench
from nglod_b200.lib.tracer import SphereTracer
from nglod_b200.lib.geoutils import _window
dev = torch.device('cuda', 0)
net, args = bench.build_and_fit(dev, print)
tr = SphereTracer(args)
torch.manual_seed(1000)
wx, wy = _window(bench.W, bench.H, "cpu"); wx, wy = wx.pin_memory(), wy.pin_memory()
n = bench.W * bench.H
out = {"depth": torch.empty(n, 1).pin_memory(), "hit": torch.empty(n, dtype=torch.bool).pin_memory(), "normal": torch.empty(n, 3).pin_memory()}
for chunks in (1, 2, 3, 4, 6):
    for streams in (1, 2, 3):
        fn = lambda: tr.trace_lookat_host(net, bench.CAM_FROM, bench.CAM_TO, bench.W, bench.H, fov=bench.FOV, window=(wx, wy), out=out, chunks=chunks, streams=streams)
        for _ in range(3): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(30): fn()
        print(f"chunks {chunks} streams {streams}: {(time.perf_counter() - t0) / 30 * 1e3:.3f} ms", flush=True)
