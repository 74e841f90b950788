"""Launch list target: three 500k-point FusedTrainer steps (ncu --metrics gpu__time_duration.sum)."""
import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from helpers import rand5_model
from nglod_b200.lib.trainer import FusedTrainer
dev = torch.device("cuda", 0)
net, _ = rand5_model(dev); net.train()
tr = FusedTrainer(net, lr=1e-3)
g = torch.Generator(device=dev).manual_seed(1)
pts = torch.rand(500000, 3, device=dev, generator=g) * 2 - 1; gts = torch.rand(500000, 1, device=dev, generator=g)
for _ in range(3): tr.step(pts, gts)
torch.cuda.synchronize()
