"""GPU-side timeline of the pipelined shade_images (mode A: trace + shade per range on alternating streams, copies on a third):
CUDA events after every piece, ms from the start of the frame."""
import copy, sys, time, torch, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from nglod_b200 import ops
from nglod_b200.lib.renderer import Renderer
from nglod_b200.lib.tracer import SphereTracer
from nglod_b200.lib.tracer.SphereTracer import _trace_lod
from nglod_b200.lib.geoutils import _window, camera_basis
dev = torch.device('cuda', 0)
net, args = bench.build_and_fit(dev, lambda m: None)
a2 = copy.copy(args); a2.render_res = [bench.W, bench.H]
r = Renderer(SphereTracer(a2), args=a2, device=dev)
r.shade_images(net, f=bench.CAM_FROM, t=bench.CAM_TO, fov=bench.FOV)
ws = r._pipe_ws; W, H = bench.W, bench.H; n = W * H; tr = r.tracer; far = r.camera_clamp[1]
view, lod = net.net_view(), _trace_lod(net); tex = r._get_matcap(dev).tex
shapes = {"x": (3, torch.float32), "hit": (1, torch.bool), "depth": (1, torch.float32), "relative_depth": (1, torch.float32),
          "normal": (3, torch.float32), "rgb": (3, torch.float32), "view": (3, torch.float32)}
host = {k: torch.empty((n, c), dtype=dt, pin_memory=True) for k, (c, dt) in shapes.items()}
cur = torch.cuda.current_stream(dev); s_out = ws["s_out"]
queue = torch.empty(16, dtype=torch.int32, device=dev)
def frame(chunks, shade_on, copy_on, reserve=0):
    marks = []
    def mark(name, stream):
        e = torch.cuda.Event(enable_timing=True); e.record(stream); marks.append((name, e))
    torch.cuda.synchronize()
    mark("start", cur)
    origin, cview, right, up = camera_basis(bench.CAM_FROM, bench.CAM_TO)
    wx, wy = _window(W, H, dev)
    ops.generate_rays(origin, cview, right, up, np.float32(np.tan(np.radians(bench.FOV / 2))), False, wx, wy, out=(ws["o"], ws["d"]))
    mark("rays", cur)
    s_out.wait_stream(cur)
    if copy_on:
        with torch.cuda.stream(s_out):
            host["view"].copy_(ws["d"], non_blocking=True); mark("copy view", s_out)
    bounds = [(n * i) // chunks for i in range(chunks + 1)]
    for i in range(chunks):
        a, b = bounds[i], bounds[i + 1]
        sc = ws["s_c"][i % 2]; sc.wait_stream(cur)
        with torch.cuda.stream(sc):
            ops.sphere_trace(view, lod, ws["o"][a:b], ws["d"][a:b], num_steps=tr.num_steps, step_size=tr.step_size, min_dis=tr.min_dis, far=far,
                             out=(ws["x"][a:b], ws["depth"][a:b], ws["hit"][a:b], ws["normal"][a:b]), queue=queue[i:i + 1],
                             max_ctas=(148 - reserve) if (reserve and i > 0) else 0)
            mark(f"trace {i}", sc)
            if shade_on:
                torch.div(torch.clamp(ws["depth"][a:b], 0.0, far), far, out=ws["relative_depth"][a:b])
                ops.shade_matcap(ws["d"][a:b], ws["normal"][a:b], ws["hit"][a:b], tex, out=ws["rgb"][a:b])
                mark(f"shade {i}", sc)
            ev = torch.cuda.Event(); ev.record(sc)
        s_out.wait_event(ev)
        if copy_on:
            with torch.cuda.stream(s_out):
                for k in ("x", "hit", "depth", "relative_depth", "normal", "rgb"):
                    host[k][a:b].copy_(ws[k][a:b].reshape(b - a, -1), non_blocking=True)
                mark(f"copies {i}", s_out)
    torch.cuda.synchronize()
    return [(nm, marks[0][1].elapsed_time(e)) for nm, e in marks[1:]]
for chunks, shade_on, copy_on, reserve in ((3, True, True, 0), (3, True, True, 8), (3, True, False, 8), (4, True, True, 8)):
    for _ in range(3): frame(chunks, shade_on, copy_on, reserve)
    tl = frame(chunks, shade_on, copy_on, reserve)
    print(f"chunks {chunks} shade {shade_on} copies {copy_on} reserve {reserve}: " + ", ".join(f"{nm} {t:.3f}" for nm, t in tl))
