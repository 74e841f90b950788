"""720p bench frame + 2^20 forward/backward timing for the current build (L2 flushed, CUDA events, median of 15)."""
import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from nglod_b200 import ops
from nglod_b200.lib.tracer import SphereTracer
dev = torch.device('cuda', 0)
net, args = bench.build_and_fit(dev, lambda m: None)
ray_o, ray_d = bench.make_rays(dev)
tracer = SphereTracer(args)
view = net.net_view()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, it=15):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(it):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))
rb = tracer(net, ray_o, ray_d)
g = torch.Generator(device=dev).manual_seed(1)
xq = torch.rand(1 << 20, 3, device=dev, generator=g) * 2 - 1
gq = torch.rand(1 << 20, device=dev, generator=g)
grid_grads = [torch.zeros_like(f.fm.data, memory_format=torch.preserve_format) for f in net.features]
dec_grad = tuple(torch.zeros_like(p) for p in net.decoder_params(bench.LOD))
scratch = net.summed_grad_scratch()
print(f"trace 720p: {timed(lambda: tracer(net, ray_o, ray_d)):.4f} ms  hits {int(rb.hit.sum())}  checksum {float(rb.depth.double().sum()):.6f} "
      f"| forward 2^20: {timed(lambda: ops.sdf_forward(view, bench.LOD, xq)) * 1e3:.1f} us "
      f"| backward 2^20: {timed(lambda: ops.sdf_backward(view, bench.LOD, xq, gq, grid_grads, dec_grad, summed_scratch=scratch)) * 1e3:.1f} us", flush=True)
