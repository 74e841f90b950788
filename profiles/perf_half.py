"""fp32 vs fp16 grid storage: forward and 720p tracer timing on the bench's fitted torus model, plus how far the
fp16-storage frame is from the fp32-storage frame (hit mask, depth, normals)."""
import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from nglod_b200 import ops
from nglod_b200.lib.tracer import SphereTracer
dev = torch.device('cuda', 0)
net, args = bench.build_and_fit(dev, print)
ray_o, ray_d = bench.make_rays(dev)
g = torch.Generator(device=dev).manual_seed(1)
xq = torch.rand(1 << 20, 3, device=dev, generator=g) * 2 - 1
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tracer = SphereTracer(args)

def timeit(fn, n=20):
    for _ in range(3): flush.zero_(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.mean(ts)), float(min(ts))

res = {}
for mode, storage, summ in (("fp32", "fp32", False), ("tc", "fp32", False), ("fp32", "fp32", True), ("tc", "fp32", True), ("tc", "fp16", True)):
    net.math_mode, net.grid_storage, net.sum_lods = mode, storage, summ
    view = net.net_view()
    f_ms, f_min = timeit(lambda: ops.sdf_forward(view, 4, xq))
    t_ms, t_min = timeit(lambda: tracer(net, ray_o, ray_d))
    rb = tracer(net, ray_o, ray_d)
    res[(mode, storage, summ)] = (rb, ops.sdf_forward(view, 4, xq))
    print(f"{mode:5s} grids {storage} summed={summ!s:5s}: forward {f_ms:.4f} ms ({(1<<20)/f_ms*1e3:.3e} q/s, min {f_min:.4f}); "
          f"tracer {t_ms:.4f} ms ({ray_o.shape[0]/t_ms*1e3:.3e} rays/s, {1e3/t_ms:.0f} fps, min {t_min:.4f})")
def cmp(ka, kb, tag):
    a, da = res[ka]; b, db = res[kb]
    both = a.hit & b.hit
    print(tag, "sdf max abs diff", float((da - db).abs().max()))
    print(tag, "tracer: hit mismatches", int((a.hit != b.hit).sum()), "of", int(a.hit.sum()), "hits;",
              "depth diff max", float((a.depth - b.depth).abs()[both].max()), "mean", float((a.depth - b.depth).abs()[both].mean()),
          "; normal diff max", float((a.normal - b.normal).abs()[both].max()), " >1e-3:", int(((a.normal - b.normal).abs().max(dim=1)[0][both] > 1e-3).sum()))
cmp(("tc", "fp32", False), ("tc", "fp32", True), "[summed-vs-per-LOD, fp32]")
cmp(("tc", "fp32", True), ("tc", "fp16", True), "[fp16-vs-fp32 summed]")
