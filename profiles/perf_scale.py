"""Forward time vs N (fixed cost vs slope) for the default inference path, random-init 5-LOD model."""
import sys, torch, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from nglod_b200 import ops
from helpers import rand5_model
dev = torch.device('cuda', 0)
net, args = rand5_model(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
g = torch.Generator(device=dev).manual_seed(1)
for storage in ("fp32", "fp16"):
    net.grid_storage = storage
    view = net.net_view()
    prev = None
    for lg in (10, 16, 18, 19, 20, 21, 22, 23):
        n = 1 << lg
        xq = torch.rand(n, 3, device=dev, generator=g) * 2 - 1
        for _ in range(3): ops.sdf_forward(view, 4, xq)
        ts = []
        for _ in range(10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.sdf_forward(view, 4, xq); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        t = float(np.median(ts))
        print(f"{storage} N=2^{lg}: {t*1e3:8.1f} us  {n/t*1e3:.3e} q/s" + (f"  marginal {(n - prev[0])/(t - prev[1])*1e3:.3e} q/s" if prev and t > prev[1] else ""))
        prev = (n, t)
