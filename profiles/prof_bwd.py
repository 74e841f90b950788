"""ncu target: a few single-grid backward launches at 2^20 queries (lod 4)."""
import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from nglod_b200 import ops, _lib
from helpers import rand5_model
dev = torch.device('cuda', 0)
net, args = rand5_model(dev)
g = torch.Generator(device=dev).manual_seed(1)
n = 1 << 20
xq = torch.rand(n, 3, device=dev, generator=g) * 2 - 1
gq = torch.rand(n, device=dev, generator=g)
view = net.net_view(inference=False)
grid_grads = [torch.zeros_like(f.fm.data, memory_format=torch.preserve_format) for f in net.features]
scratch = [torch.zeros_like(f.fm.data, memory_format=torch.preserve_format) for f in net.features]
dec_grads = [tuple(torch.zeros_like(p) for p in net.decoder_params(l)) for l in range(5)]
for _ in range(3):
    ops.sdf_backward(view, 4, xq, gq, grid_grads, dec_grads[4], summed_scratch=scratch)
torch.cuda.synchronize()
