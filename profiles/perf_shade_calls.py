import sys, time, torch
sys.path.insert(0, '/root/repo')
import bench
from nglod_b200.lib.tracer import SphereTracer
from nglod_b200.lib.renderer import Renderer
dev = torch.device('cuda', 0)
net, args = bench.build_and_fit(dev, print)
r = Renderer(SphereTracer(args), args=args, device=dev)
f, t = bench.CAM_FROM, bench.CAM_TO
big = torch.empty(1 << 23, 3, device=dev)    # like the bench's variants block
hp = [torch.empty(921600, 3).pin_memory() for _ in range(4)]
for i in range(12):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    img = r.shade_images(net, f=f, t=t, fov=30.0)
    torch.cuda.synchronize(); print(i, f"{(time.perf_counter()-t0)*1e3:.2f} ms")
st = r.shade_tensor(net, f=f, t=t, fov=30.0, mm=torch.eye(3))
for i in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c = st.cpu()
    torch.cuda.synchronize(); print("cpu()", i, f"{(time.perf_counter()-t0)*1e3:.2f} ms")
