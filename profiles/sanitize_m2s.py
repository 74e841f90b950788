import sys, torch
sys.path.insert(0, "/root/repo")
from nglod_b200 import ops
from nglod_b200.lib.torchgp import icosphere
V, F = icosphere(3)
tri = V.cuda()[F.cuda()].contiguous()
for n in (100, 5000, 20000):
    p = torch.rand(n, 3, device="cuda") * 2 - 1
    d = ops.mesh2sdf_gpu(p, tri)[0]
torch.cuda.synchronize(); print("done", tri.shape, float(d.min()), float(d.max()))
