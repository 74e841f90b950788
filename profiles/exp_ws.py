"""Forward kernel timing + identity check for the current build (used by the warp-specialisation sweep)."""
import sys, torch, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from nglod_b200 import ops
from helpers import rand5_model
dev = torch.device('cuda', 0)
net, args = rand5_model(dev)
g = torch.Generator(device=dev).manual_seed(1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, it=20, do_flush=True):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(it):
        if do_flush: flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), min(ts)
res = {}
for storage in ("fp32", "fp16"):
    net.grid_storage = storage
    net.math_mode = "fp32"; vf = net.net_view()
    net.math_mode = "tc"; vt = net.net_view()
    for logn in (10, 14, 17, 20, 20.5, 23):
        n = int(2 ** logn) + (3 if logn == 20.5 else 0)
        xq = torch.rand(n, 3, device=dev, generator=g) * 2 - 1
        d_tc = ops.sdf_forward(vt, 4, xq)
        d_32 = ops.sdf_forward(vf, 4, xq)
        err = (d_tc - d_32).abs().max().item()
        med, mn = timed(lambda: ops.sdf_forward(vt, 4, xq))
        med2, _ = timed(lambda: ops.sdf_forward(vt, 4, xq), do_flush=False)
        print(f"{storage} n=2^{logn}: {med*1e3:8.1f} us (min {mn*1e3:.1f}) = {n/med/1e6:6.2f} Gq/s | no flush {med2*1e3:8.1f} us = {n/med2/1e6:6.2f} Gq/s | max|tc-fp32| {err:.2e}", flush=True)
