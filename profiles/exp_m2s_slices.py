"""[historical: the NGLOD_M2S_DIST_SLICES knob existed only for this sweep; the result is recorded in profiles/README.md]
Distance-kernel slices sweep (NGLOD_M2S_DIST_SLICES), fixed seed, 500 k points."""
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
for name, (V, F) in (("torus", torus(0.6, 0.25, 128, 64)), ("ico5", icosphere(5)), ("ico6", icosphere(6))):
    V, F = normalize(V.to(dev), F.to(dev)); tri = V[F].contiguous()
    for seed in (0, 1):
        torch.manual_seed(seed)
        pts = point_sample(V, F, ["rand", "near", "near", "trace", "trace"], 100000)
        row = []
        for sl in (4, 8, 16, 32, 64):
            os.environ["NGLOD_M2S_DIST_SLICES"] = str(sl)
            row.append("%d: %.2f" % (sl, t(lambda: ops.mesh2sdf_gpu(pts, tri))))
        print(name, tri.shape[0], "seed", seed, " | ".join(row), "| rand only:", end=" ")
        r = pts[:100000].contiguous(); row = []
        for sl in (8, 32):
            os.environ["NGLOD_M2S_DIST_SLICES"] = str(sl)
            row.append("%d: %.2f" % (sl, t(lambda: ops.mesh2sdf_gpu(r, tri))))
        print(" | ".join(row))
