import sys, time, torch
sys.path.insert(0, '/root/repo')
import bench
from nglod_b200 import ops
from nglod_b200.lib.tracer import SphereTracer
from nglod_b200.lib.renderer import Renderer
dev = torch.device('cuda', 0)
net, args = bench.build_and_fit(dev, print)
r = Renderer(SphereTracer(args), args=args, device=dev)
f, t = bench.CAM_FROM, bench.CAM_TO
def T(name, fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): out = fn()
    torch.cuda.synchronize(); print(f"{name:44s} {(time.perf_counter()-t0)/n*1e3:8.3f} ms"); return out
mm = torch.eye(3)
rb = T("render_lookat(mm=eye)", lambda: r.render_lookat(net, f=f, t=t, fov=30.0, mm=mm))
rb = T("render_lookat(mm=None)", lambda: r.render_lookat(net, f=f, t=t, fov=30.0, mm=None))
view = rb.view
mc = r._get_matcap(view.device)
T("_get_matcap", lambda: r._get_matcap(view.device))
T("view mm", lambda: torch.mm(view.reshape(-1, 3), mm.to(view.device).transpose(1, 0)).reshape(1280, 720, 3))
rb.normal = rb.normal.contiguous()
T("shade_matcap", lambda: ops.shade_matcap(view, rb.normal, rb.hit, mc.tex))
st = T("shade_tensor(mm=eye)", lambda: r.shade_tensor(net, f=f, t=t, fov=30.0, mm=mm))
st = T("shade_tensor(mm=None)", lambda: r.shade_tensor(net, f=f, t=t, fov=30.0, mm=None))
T("rb.cpu()", lambda: st.cpu())
keep = []
T("rb.cpu() keeping results", lambda: keep.append(st.cpu()), n=5)
T("shade_images", lambda: r.shade_images(net, f=f, t=t, fov=30.0))
