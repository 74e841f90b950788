import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from helpers import rand5_model
from nglod_b200.lib.trainer import FusedTrainer
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
pts = torch.rand(512, 3, device=dev, generator=g) * 2 - 1; gts = torch.rand(512, 1, device=dev, generator=g)
net, _ = rand5_model(dev); net.train()
tr = FusedTrainer(net, lr=1e-3, use_graph=False)
for _ in range(4): tr.step(pts, gts)
torch.cuda.synchronize()
