"""mesh2sdf per-kernel times vs batch size (run under ncu --metrics gpu__time_duration.sum; one call per size)."""
import sys, torch
sys.path.insert(0, '/root/repo')
from nglod_b200 import ops
from nglod_b200.lib.torchgp import torus, point_sample, normalize
dev = 'cuda'
V, F = normalize(*[t.to(dev) for t in torus(0.6, 0.25, 128, 64)])
tri = V[F].contiguous()
torch.manual_seed(0)
for n in (62500, 125000, 250000, 500000):
    pts = point_sample(V, F, ["rand", "near", "near", "trace", "trace"], n // 5)
    for _ in range(2): ops.mesh2sdf_gpu(pts, tri)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): ops.mesh2sdf_gpu(pts, tri)
    b.record(); torch.cuda.synchronize()
    print(f"n={n}: {a.elapsed_time(b) / 5:.3f} ms per call", flush=True)
