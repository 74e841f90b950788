"""Experiment: the tracer writes depth / hit / normal straight into PINNED HOST buffers (zero copy over PCIe, posted writes
as the rays retire) instead of device buffers + cudaMemcpyAsync afterwards.  One launch per frame."""
import ctypes, sys, time, torch, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from nglod_b200 import ops, _lib
from nglod_b200.lib.tracer import SphereTracer
lib = _lib.load()
dev = torch.device('cuda', 0)
net, args = bench.build_and_fit(dev, lambda m: None)
ray_o, ray_d = bench.make_rays(dev)
n = ray_o.shape[0]
view = net.net_view()
tracer = SphereTracer(args)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
P = lambda t: ctypes.c_void_p(t.data_ptr())
opts = _lib.TraceOpts(256, 1, 1.0, 0.0003, 10.0, 1.0 / (64.0 * 3.0))
queue = torch.empty(1, dtype=torch.int32, device=dev)
x_d = torch.empty(n, 3, device=dev)
def run(depth, hit, normal, x):
    _lib.check(lib.nglod_sphere_trace(ctypes.byref(view.struct), bench.LOD, P(ray_o), P(ray_d), n, ctypes.byref(opts), P(x), P(depth),
                                      P(hit), P(normal), P(queue), ctypes.c_void_p(0),
                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "trace")
ref = ops.sphere_trace(view, bench.LOD, ray_o, ray_d)
for tag, mk in (("device buffers", lambda s, dt: torch.empty(s, dtype=dt, device=dev)),
                ("pinned host buffers (zero copy)", lambda s, dt: torch.empty(s, dtype=dt).pin_memory())):
    depth, hit, normal = mk((n, 1), torch.float32), mk((n,), torch.bool), mk((n, 3), torch.float32)
    for _ in range(3): run(depth, hit, normal, x_d)
    torch.cuda.synchronize()
    ts, ws = [], []
    for _ in range(10):
        flush.zero_(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record(); run(depth, hit, normal, x_d); b.record(); torch.cuda.synchronize()
        ws.append((time.perf_counter() - t0) * 1e3); ts.append(a.elapsed_time(b))
    same = torch.equal(depth.to(dev), ref[1]) and torch.equal(hit.to(dev), ref[2]) and torch.equal(normal.to(dev), ref[3])
    print(f"{tag:34s}: kernel {np.median(ts):.3f} ms, wall incl. sync {np.median(ws):.3f} ms, identical {same}", flush=True)
# x too
x_h = torch.empty(n, 3).pin_memory()
depth, hit, normal = (torch.empty((n, 1)).pin_memory(), torch.empty(n, dtype=torch.bool).pin_memory(), torch.empty(n, 3).pin_memory())
for _ in range(3): run(depth, hit, normal, x_h)
torch.cuda.synchronize(); ts = []
for _ in range(10):
    flush.zero_(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(depth, hit, normal, x_h); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
print(f"all four outputs pinned (29 B/ray): kernel {np.median(ts):.3f} ms")
