import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from nglod_b200 import ops
sys.path.insert(0, "/root/repo/tests")
from helpers import rand5_model
sys.path.insert(0,'/root/repo/tests')
dev = torch.device('cuda', 0)
net, args = rand5_model(dev)
g = torch.Generator(device=dev).manual_seed(1)
xq = torch.rand(1 << 20, 3, device=dev, generator=g) * 2 - 1
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for mode in ("fp32", "tc"):
    net.math_mode = mode
    view = net.net_view()
    for _ in range(3): ops.sdf_forward(view, 4, xq)
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.sdf_forward(view, 4, xq); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(mode, f"{np.mean(ts):.4f} ms  -> {(1<<20)/np.mean(ts)*1e3:.3e} q/s  (min {min(ts):.4f})")
