import sys, torch, time
sys.path.insert(0, "/root/repo")
from nglod_b200 import ops
from nglod_b200.lib.torchgp import torus, normalize
V, F = normalize(*[t.cuda() for t in torus(0.6, 0.25, 128, 64)])
tri = V[F].contiguous()
g = torch.Generator(device="cuda").manual_seed(0)
for n in (1000, 5000, 20000, 62500):
    p = torch.rand(n, 3, device="cuda", generator=g) * 2 - 1
    for rep in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); d = ops.mesh2sdf_gpu(p, tri)[0]; b.record(); torch.cuda.synchronize()
        print(n, rep, "gpu", round(a.elapsed_time(b), 3), "ms  wall", round((time.perf_counter() - t0) * 1e3, 3))
