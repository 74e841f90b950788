"""Small shapes through every kernel, for compute-sanitizer (memcheck / racecheck / initcheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch
from helpers import make_args
from nglod_b200 import ops
from nglod_b200.lib.models import OctreeSDF
from nglod_b200.lib.tracer import SphereTracer
from nglod_b200.lib.trainer import FusedTrainer
from nglod_b200.lib import spc as S
from nglod_b200.lib.torchgp import icosphere
from nglod_b200.lib.geoutils import look_at
dev = "cuda"
args = make_args(["--num-lods", "3"])
torch.manual_seed(0)
net = OctreeSDF(args).to(dev)
x = torch.rand(777, 3, device=dev) * 2.2 - 1.1
for mode, storage, summ in (("fp32", "fp32", True), ("tc", "fp32", True), ("tc", "fp16", True), ("tc", "fp32", False), ("fp32", "fp32", False)):
    net.math_mode, net.grid_storage, net.sum_lods = mode, storage, summ       # every gather / backward variant
    net.lod = 2
    with torch.no_grad():
        net.sdf(x, lod=2); net.features[1](x)
    d = net.sdf(x.clone().requires_grad_(True), lod=1); d.sum().backward()
    ro, rd = look_at([-2.8, 2.8, -2.8], [0, 0, 0], 48, 27, mode="persp", fov=30.0, device=dev)
    SphereTracer(args)(net, ro, rd)
    SphereTracer(args).trace_host(net, ro.cpu().pin_memory(), rd.cpu().pin_memory(), chunks=3)
    from nglod_b200.lib.diffutils import gradient
    gradient(x, net, method="finitediff")
V, F = icosphere(2)
pts = torch.rand(5000, 3, device=dev) * 2 - 1
ops.mesh2sdf_gpu(pts, V.to(dev)[F.to(dev)].contiguous()); ops.mesh2sdf_gpu(pts[:100], V.to(dev)[F.to(dev)].contiguous())
net.sum_lods = True
tr = FusedTrainer(net); tr.step(x, torch.rand(777, 1, device=dev))
net.sum_lods = False
tr.step(x, torch.rand(777, 1, device=dev))
net.sum_lods = True
octree = S.mesh_to_octree(V.to(dev), F.to(dev), 4, num_samples=1 << 16)
sp = S.SparseOctreeSDF(net, S.SPC(octree))
for mode in ("fp32", "tc"):
    sp.math_mode = mode
    sp.trace(ro, rd, 2)
ops.aabb(ro, rd)
torch.cuda.synchronize()
print("sanitize target done")
