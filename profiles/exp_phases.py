"""Where a tracer round goes: per-phase clock64() sums (build with -DNGLOD_TRACE_TIMING; experiment only)."""
import sys, torch
sys.path.insert(0, '/root/repo')
import bench
from nglod_b200 import ops
dev = torch.device('cuda', 0)
net, args = bench.build_and_fit(dev, print)
ray_o, ray_d = bench.make_rays(dev)
for storage in ("fp32", "fp16"):
    net.grid_storage = storage
    view = net.net_view()
    for _ in range(2): ops.sphere_trace(view, 4, ray_o, ray_d)
    stats = torch.zeros(8, dtype=torch.int64, device=dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ops.sphere_trace(view, 4, ray_o, ray_d, stats=stats); b.record(); torch.cuda.synchronize()
    s = stats.cpu().tolist()
    names = ["refill", "gather", "group barrier", "mma wait", "epilogue", "state machine"]
    tot = sum(s[2:8])
    print(f"{storage}: {a.elapsed_time(b):.3f} ms, evals {s[0]}, per-warp-cycle shares:")
    for n, v in zip(names, s[2:8]): print(f"   {n:14s} {100.0 * v / tot:5.1f} %   ({v / (148 * 16):.0f} cycles per warp)")
