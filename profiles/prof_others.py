"""Profiling target for the non-headline kernels: backward, mesh2sdf, SPC traversal / in-voxel tracer."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch
from helpers import rand5_model
from nglod_b200 import ops
from nglod_b200.lib import spc as S
from nglod_b200.lib.torchgp import torus, normalize, point_sample
from nglod_b200.lib.geoutils import look_at
dev = torch.device("cuda", 0)
net, args = rand5_model(dev)
view = net.net_view()
g = torch.Generator(device=dev).manual_seed(1)
xq = torch.rand(1 << 20, 3, device=dev, generator=g) * 2 - 1
gq = torch.rand(1 << 20, device=dev, generator=g)
grid_grads = [torch.zeros_like(f.fm.data, memory_format=torch.preserve_format) for f in net.features]
dec_grad = tuple(torch.zeros_like(p) for p in net.decoder_params(4))
for _ in range(2):
    ops.sdf_backward(view, 4, xq, gq, grid_grads, dec_grad)
V, F = normalize(*[t.to(dev) for t in torus(0.6, 0.25, 128, 64)])
pts = point_sample(V, F, ["rand", "near", "trace"], 40000)
for _ in range(2):
    ops.mesh2sdf_gpu(pts, V[F].contiguous())
octree = S.mesh_to_octree(V, F, 6, num_samples=1 << 21)
spc = S.SPC(octree)
sp = S.SparseOctreeSDF(net, spc)
ro, rd = look_at([-2.8, 2.8, -2.8], [0, 0, 0], 1920, 1080, mode="persp", fov=30.0, device=dev)
for _ in range(2):
    sp.trace(ro, rd, 4)
torch.cuda.synchronize()
