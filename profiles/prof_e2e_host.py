"""Host-side cost of SphereTracer.trace_lookat_host(packed=True), piece by piece (wall per frame, back to back)."""
import sys, time, torch, numpy as np
import torch.nn.functional as F
sys.path.insert(0, '/root/repo')
import bench
from nglod_b200 import ops
from nglod_b200.lib.tracer import SphereTracer
from nglod_b200.lib.tracer.SphereTracer import _trace_lod
from nglod_b200.lib.tracer.RenderBuffer import RenderBuffer
from nglod_b200.lib.geoutils import _window
dev = torch.device('cuda', 0)
net, args = bench.build_and_fit(dev, lambda m: None)
tracer = SphereTracer(args)
W, H = bench.W, bench.H
n = W * H
torch.manual_seed(1000)
wx, wy = _window(W, H, "cpu"); wx, wy = wx.pin_memory(), wy.pin_memory()
cam = bench.camera_from(0.0)
out = {}
def full(): return tracer.trace_lookat_host(net, cam, bench.CAM_TO, W, H, fov=bench.FOV, window=(wx, wy), out=out, packed=True)
def basis():
    origin = torch.tensor(list(cam), dtype=torch.float32)
    view = F.normalize(torch.tensor(list(bench.CAM_TO), dtype=torch.float32) - origin, dim=0)
    right = F.normalize(torch.linalg.cross(view, torch.tensor([0.0, 1.0, 0.0])), dim=0)
    up = F.normalize(torch.linalg.cross(right, view), dim=0)
    return origin.tolist(), view.tolist(), right.tolist(), up.tolist()
B = basis()
tan = np.float32(np.tan(np.radians(bench.FOV / 2)))
wsd = torch.empty(6 * n + W + H, device=dev)
p16, hit_h = torch.empty(n, 4).pin_memory(), torch.empty(n, dtype=torch.bool).pin_memory()
q = torch.empty(1, dtype=torch.int32, device=dev)
nv = net.net_view()
def core(b=B, v=nv):
    ops.sphere_trace_camera(v, bench.LOD, b[0], b[1], b[2], b[3], tan, False, wx, wy, wsd, p16, hit=hit_h, queue=q,
                            num_steps=tracer.num_steps, step_size=tracer.step_size, min_dis=tracer.min_dis, far=tracer.camera_clamp[1])
    torch.cuda.current_stream(dev).synchronize()
def with_basis(): core(b=basis())
def with_view(): core(v=net.net_view())
def with_dev():
    d = next(net.parameters()).device
    with torch.cuda.device(d): core()
def with_rb():
    core(); return RenderBuffer(**ops.unpack_trace(p16, hit_h))
def wall(fn, it=100):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(it): fn()
    return (time.perf_counter() - t0) / it * 1e3
for rep in range(2):
    for name, fn in (("trace_lookat_host", full), ("core call + sync", core), ("+ camera basis (torch)", with_basis),
                     ("+ net_view()", with_view), ("+ device lookup / guard", with_dev), ("+ RenderBuffer views", with_rb)):
        print("%-28s %.3f ms" % (name, wall(fn)))
def host_only(fn, it=300):
    t0 = time.perf_counter()
    for _ in range(it): fn()
    return (time.perf_counter() - t0) / it * 1e6
print("host us: basis %.1f, net_view %.1f, _trace_lod %.1f, params-device %.1f, unpack+RenderBuffer %.1f" % (
    host_only(basis), host_only(net.net_view), host_only(lambda: _trace_lod(net)), host_only(lambda: next(net.parameters()).device),
    host_only(lambda: RenderBuffer(**ops.unpack_trace(p16, hit_h)))))
