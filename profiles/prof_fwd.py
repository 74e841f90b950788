"""ncu target: the forward kernel alone on the random-init 5-LOD model (2^20 lod-4 queries, tensor-core mode)."""
import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from nglod_b200 import ops
from helpers import rand5_model
dev = torch.device('cuda', 0)
net, args = rand5_model(dev)
net.math_mode = "tc"
v = net.net_view()
g = torch.Generator(device=dev).manual_seed(1)
xq = torch.rand(1 << 20, 3, device=dev, generator=g) * 2 - 1
for _ in range(4):
    ops.sdf_forward(v, 4, xq)
torch.cuda.synchronize()
