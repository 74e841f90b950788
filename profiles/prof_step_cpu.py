import sys, time, torch, cProfile, pstats
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from helpers import rand5_model
from nglod_b200.lib.trainer import FusedTrainer
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
B = 512
pts = torch.rand(B, 3, device=dev, generator=g) * 2 - 1; gts = torch.rand(B, 1, device=dev, generator=g)
net, _ = rand5_model(dev); net.train()
tr = FusedTrainer(net, lr=1e-3)
for _ in range(10): tr.step(pts, gts)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(300): tr.step(pts, gts)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
