"""Config 3 at 8 GPUs is 62 500 points per rank: host enqueue time against GPU time of sample + label and of the fused step."""
import copy, cProfile, io, pstats, sys, time, torch, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from nglod_b200 import ops
from nglod_b200.lib.trainer import FusedTrainer
from nglod_b200.lib.torchgp import torus, normalize, point_sample
dev = torch.device('cuda', 0)
net, args = bench.build_and_fit(dev, lambda m: None)
V, F = normalize(*[t.to(dev) for t in torus(0.6, 0.25, 128, 64)])
tri = V[F].contiguous()
modes = ["rand", "near", "near", "trace", "trace"]
for per_rank in (62500, 500000):
    def make_batch():
        pts = point_sample(V, F, modes, per_rank // 5)
        return pts, ops.mesh2sdf_gpu(pts, tri)[0].unsqueeze(1)
    tnet = copy.deepcopy(net); tnet.train()
    trainer = FusedTrainer(tnet, lr=1e-3)
    pts, gts = make_batch()
    step = lambda: trainer.step(pts, gts, global_batch=per_rank)
    for name, fn in (("sample+label", make_batch), ("trainer.step", step)):
        for _ in range(5): fn()
        torch.cuda.synchronize()
        host, tot = [], []
        for _ in range(20):
            torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
            host.append((t1 - t0) * 1e3); tot.append((t2 - t0) * 1e3)
        print(f"{per_rank:7d} {name:13s}: host enqueue {np.median(host):.3f} ms, until GPU done {np.median(tot):.3f} ms")
    if per_rank == 62500:
        pr = cProfile.Profile(); pr.enable()
        for _ in range(100): make_batch()
        torch.cuda.synchronize(); pr.disable()
        s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18); print(s.getvalue()[:4200])
