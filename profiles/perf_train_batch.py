"""FusedTrainer step time vs batch size, through the prefix-summed grids (rebuilt every step) or the per-LOD gather."""
import sys, time, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from helpers import rand5_model
from nglod_b200.lib.trainer import FusedTrainer
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
for B in (512, 4096, 16384, 65536, 500000):
    pts = torch.rand(B, 3, device=dev, generator=g) * 2 - 1; gts = torch.rand(B, 1, device=dev, generator=g)
    row = []
    for summ in (True, False):
        net, _ = rand5_model(dev); net.sum_lods = summ; net.train()
        tr = FusedTrainer(net, lr=1e-3)
        for _ in range(5): tr.step(pts, gts)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        n = 50 if B <= 65536 else 10
        for _ in range(n): tr.step(pts, gts)
        torch.cuda.synchronize(); row.append((time.perf_counter() - t0) / n * 1e3)
    print(f"B={B:7d}: summed {row[0]:.3f} ms   per-LOD {row[1]:.3f} ms")
