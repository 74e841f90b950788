"""One warm and one profiled mesh2sdf call (500 k points: rand / near / near / trace / trace, torus 16 384 triangles) for the ncu launch list."""
import sys, torch
sys.path.insert(0, '/root/repo')
from nglod_b200 import ops
from nglod_b200.lib.torchgp import torus, icosphere, point_sample, normalize
dev = 'cuda'
which = sys.argv[1] if len(sys.argv) > 1 else "torus"
V, F = torus(0.6, 0.25, 128, 64) if which == "torus" else icosphere(5)
V, F = normalize(V.to(dev), F.to(dev))
tri = V[F].contiguous()
pts = point_sample(V, F, ["rand", "near", "near", "trace", "trace"], 100000)
if len(sys.argv) > 2:
    pts = pts[:int(sys.argv[2])].contiguous()
for _ in range(2):
    ops.mesh2sdf_gpu(pts, tri)
torch.cuda.synchronize()
