"""compute-sanitizer target for the kernels added after the first sanitizer pass: summed-grid build (direct and level by
level), fp16 pack, single-grid forward / tracer (fp32 + fp16 lines), gen-2 backward + restriction cascade, fused train
step, sparse tracer on prefix-summed corner rows, real-time loop."""
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
from nglod_b200.app import realtime
dev = "cuda"
args = make_args(["--num-lods", "3"])
torch.manual_seed(0)
net = OctreeSDF(args).to(dev)
x = torch.rand(1500, 3, device=dev) * 2.2 - 1.1
ro, rd = look_at([-2.8, 2.8, -2.8], [0, 0, 0], 48, 27, mode="persp", fov=30.0, device=dev)
for storage in ("fp32", "fp16"):
    net.grid_storage = storage
    net.mark_grids_dirty()
    with torch.no_grad():
        net.sdf(x, lod=2); net.sdf(x[:33], lod=0)
    net.lod = 2
    SphereTracer(args)(net, ro, rd)
net.grid_storage = "fp32"
d = net.sdf(x, lod=2); ((d - 0.1) ** 2).mean().backward()
d = net.sdf(x[:700], lod=1); d.sum().backward()
tr = FusedTrainer(net); tr.step(x, torch.rand(1500, 1, device=dev)); tr.step(x[:513], torch.rand(513, 1, device=dev))
# tcgen05 backward, single-grid flavour with several tiles per CTA (stage re-use, barrier parities), private scatter copies
# of the 4^3 / 8^3 heads, fold kernel, octree summed-grid build, unrolled restriction
xb = torch.rand(60001, 3, device=dev) * 2 - 1
tr2 = FusedTrainer(net, summed_min_batch=0, use_graph=False)
tr2.step(xb, torch.rand(60001, 1, device=dev)); tr2.step(xb[:40000], torch.rand(40000, 1, device=dev))
net.sum_lods = False                      # per-LOD flavour, several tiles per CTA
d = net.sdf(xb[:30000], lod=2); d.sum().backward()
net.sum_lods = True
V, F = icosphere(2)
sp = S.SparseOctreeSDF(net, S.SPC(S.mesh_to_octree(V.to(dev), F.to(dev), 4, num_samples=1 << 16)))
sp.trace(ro, rd, 2)
realtime.run(net, 48, 27, frames=2, lod=2)
nspc = S.NeuralSPC(sp.spc, num_lods=3, base_lod=2)
lp = sp.spc.level_points(4)[:, :3].float()
pidx = torch.randint(0, lp.shape[0], (700,), device=dev)
xs = (lp[pidx] + torch.rand(700, 3, device=dev)) / 16 * 2 - 1
nspc.sdf(xs, 2, pidx).sum().backward()
pidx2 = torch.randint(0, lp.shape[0], (30000,), device=dev)
xs2 = (lp[pidx2] + torch.rand(30000, 3, device=dev)) / 16 * 2 - 1
pidx2[::7] = -1                           # inert rows
nspc.sdf(xs2, 2, pidx2).sum().backward()
torch.cuda.synchronize()
print("sanitize target done")
