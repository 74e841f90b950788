"""mesh2sdf through compute-sanitizer: the brute-force walk (unsorted / sorted / sliced) and the large-batch path
(distance hierarchy + projected-bin stabs), with duplicated and zero-area triangles and points outside the box."""
import sys, torch
sys.path.insert(0, "/root/repo")
from nglod_b200 import ops
from nglod_b200.lib.torchgp import icosphere
V, F = icosphere(3)
tri = V.cuda()[F.cuda()].contiguous()
tri = torch.cat([tri, tri[:40], tri[:20, :1].expand(-1, 3, -1)], 0).contiguous()
for n in (100, 5000, 20000, 40001):
    p = torch.rand(n, 3, device="cuda") * 2.6 - 1.3
    d = ops.mesh2sdf_gpu(p, tri)[0]
torch.cuda.synchronize(); print("done", tri.shape, float(d.min()), float(d.max()))
# sampler kernel + cumulative-area table, and the training step whose level-0 scatter goes through shared memory
sys.path.insert(0, "/root/repo/tests")
from helpers import rand5_model
from nglod_b200.lib.torchgp import point_sample, sample_surface
from nglod_b200.lib.trainer import FusedTrainer
Vc, Fc = V.cuda(), F.cuda()
p = point_sample(Vc, Fc, ["rand", "near", "near", "trace", "trace"], 3001)
s, nrm = sample_surface(Vc, Fc, 777)
net, _ = rand5_model(torch.device("cuda", 0)); net.train()
tr = FusedTrainer(net, lr=1e-3, use_graph=False)
x = torch.rand(70001, 3, device="cuda") * 2 - 1
tr.step(x, torch.rand(70001, 1, device="cuda"))
torch.cuda.synchronize(); print("sampler + train step done", p.shape, s.shape, nrm.shape)
