"""ncu target: the one-pass SPC traversal of a 1920x1080 frame over the level-7 octree of the torus."""
import sys, torch
sys.path.insert(0, '/root/repo')
import bench
from nglod_b200.lib import spc as S
from nglod_b200.lib.geoutils import look_at
from nglod_b200.lib.torchgp import torus, normalize
dev = torch.device('cuda', 0)
V, F = normalize(*[t.to(dev) for t in torus(0.6, 0.25, 128, 64)])
torch.manual_seed(77)
spc = S.SPC(S.mesh_to_octree(V, F, 7, num_samples=1 << 22))
torch.manual_seed(5)
ro, rd = look_at(bench.CAM_FROM, bench.CAM_TO, 1920, 1080, mode="persp", fov=bench.FOV, device=dev)
for _ in range(3):
    S._raytrace_runs(spc, ro, rd, 7)
torch.cuda.synchronize()
