"""mesh2sdf timed with a synchronise between calls (as a training loop that consumes the labels would)."""
import sys, torch
sys.path.insert(0, '/root/repo')
from nglod_b200 import ops
from nglod_b200.lib.torchgp import torus, point_sample, normalize
dev = 'cuda'
V, F = normalize(*[t.to(dev) for t in torus(0.6, 0.25, 128, 64)])
tri = V[F].contiguous()
ts = []
for it in range(8):
    pts = point_sample(V, F, ["rand", "near", "near", "trace", "trace"], 100000)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); d = ops.mesh2sdf_gpu(pts, tri)[0]; b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print("mesh2sdf 500k x 16384, sync between calls, ms:", " ".join("%.2f" % t for t in ts))
ts = []
for it in range(8):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); pts = point_sample(V, F, ["rand", "near", "near", "trace", "trace"], 100000); d = ops.mesh2sdf_gpu(pts, tri)[0]; b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print("sample + label, ms:", " ".join("%.2f" % t for t in ts))
