import sys, time, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from helpers import fit3_model, make_args
import numpy as np
from nglod_b200.lib.tracer import SphereTracer
from nglod_b200.lib.renderer import Renderer
from nglod_b200.lib.geoutils import look_at, spherical_envmap
fit3 = dict(np.load('/root/repo/tests/golden/fit3.npz'))
net, _ = fit3_model(fit3, 'cuda'); net.lod = 2
args = make_args(["--num-lods", "3", "--render-res", "1280", "720"])
r = Renderer(SphereTracer(args), args=args, device='cuda')
def T(name, fn, n=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): out = fn()
    torch.cuda.synchronize(); print(f"{name:28s} {(time.perf_counter()-t0)/n*1e3:8.3f} ms"); return out
f, t = args.camera_origin, args.camera_lookat
ro, rd = T("look_at", lambda: look_at(f, t, 1280, 720, fov=30.0, mode='persp', device='cuda'))
rb = T("tracer", lambda: r.tracer(net, ro, rd))
rb2 = T("render (trace+reldepth+reshape)", lambda: r.render(net, ro, rd))
T("spherical_envmap", lambda: spherical_envmap(rb2.view.clone(), rb2.normal.clone()))
uv = spherical_envmap(rb2.view.clone(), rb2.normal.clone())
mc = r._get_matcap(uv.device)
T("matcap lookup", lambda: mc(uv))
st = T("shade_tensor", lambda: r.shade_tensor(net, f=f, t=t, fov=30.0, mm=torch.eye(3)))
T("rb.cpu()", lambda: st.cpu())
T("rb.cpu().transpose()", lambda: st.cpu().transpose())
T("shade_images", lambda: r.shade_images(net, f=f, t=t, fov=30.0))
