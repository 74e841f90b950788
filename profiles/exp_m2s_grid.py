"""[historical: the NGLOD_M2S_GRID_MULT knob existed only for this sweep; the result is recorded in profiles/README.md]
Projected-grid resolution sweep for the stab half (NGLOD_M2S_GRID_MULT: G^2 >= mult * #triangles)."""
import os, sys, torch
sys.path.insert(0, '/root/repo')
from nglod_b200 import ops
from nglod_b200.lib.torchgp import torus, icosphere, point_sample, normalize
dev = 'cuda'
def t(fn, it=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it
for name, (V, F) in (("torus", torus(0.6, 0.25, 128, 64)), ("ico5", icosphere(5)), ("ico3", icosphere(3)), ("ico6", icosphere(6))):
    V, F = normalize(V.to(dev), F.to(dev)); tri = V[F].contiguous()
    torch.manual_seed(0)
    pts = point_sample(V, F, ["rand", "near", "near", "trace", "trace"], 100000)
    row = []
    for mult in (1, 4, 16, 64):
        os.environ["NGLOD_M2S_GRID_MULT"] = str(mult)
        row.append("x%d: %.2f" % (mult, t(lambda: ops.mesh2sdf_gpu(pts, tri))))
    print(name, tri.shape[0], " | ".join(row))
