import sys, time, torch, traceback
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from helpers import rand5_model
from nglod_b200.lib import trainer as T
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
B = 512
pts = torch.rand(B, 3, device=dev, generator=g) * 2 - 1; gts = torch.rand(B, 1, device=dev, generator=g)
net, _ = rand5_model(dev); net.train()
tr = T.FusedTrainer(net, lr=1e-3)
# capture with the exception visible
try:
    st = {"pts": pts.clone(), "gts": gts.clone()}
    lods = tr.loss_lods
    cur = torch.cuda.current_stream(dev); side = torch.cuda.Stream(dev); side.wait_stream(cur)
    with torch.cuda.stream(side):
        for _ in range(2):
            net.mark_grids_dirty(); tr._compute_grads(st["pts"], st["gts"], B, lods)
    cur.wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        tr._compute_grads(st["pts"], st["gts"], B, lods)
    print("capture OK")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(200): graph.replay()
    torch.cuda.synchronize(); print("replay only", (time.perf_counter() - t0) / 200 * 1e3, "ms")
except Exception:
    traceback.print_exc()
for _ in range(5): tr.step(pts, gts)
print({k: (v is not False) for k, v in tr._graphs.items()})
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(200): tr.step(pts, gts)
torch.cuda.synchronize(); print("step", (time.perf_counter() - t0) / 200 * 1e3, "ms")
t0 = time.perf_counter()
for _ in range(200):
    from nglod_b200 import ops
    ops.adam_step(tr.flat, tr.flat_grad, tr.exp_avg, tr.exp_avg_sq, 5, lr=1e-3)
torch.cuda.synchronize(); print("adam only", (time.perf_counter() - t0) / 200 * 1e3, "ms")
