"""Backward / training-step timings (single-grid path): 2^20-query backward at lod 4 (kernel + cascade, kernel alone, no grid
gradient), the 500 k x 5-head fused step, and small batches.  L2 flushed before every timed launch."""
import sys, torch, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from nglod_b200 import ops, _lib
from helpers import rand5_model
dev = torch.device('cuda', 0)
net, args = rand5_model(dev)
g = torch.Generator(device=dev).manual_seed(1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, it=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(it):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
for n in (1 << 20, 500000, 65536, 4096, 512):
    xq = torch.rand(n, 3, device=dev, generator=g) * 2 - 1
    gq = torch.rand(n, device=dev, generator=g)
    view = net.net_view(inference=False)
    grid_grads = [torch.zeros_like(f.fm.data, memory_format=torch.preserve_format) for f in net.features]
    scratch = [torch.zeros_like(f.fm.data, memory_format=torch.preserve_format) for f in net.features]
    ss = torch.zeros(1 << 18, device=dev)
    dec_grads = [tuple(torch.zeros_like(p) for p in net.decoder_params(l)) for l in range(5)]
    t_full = timeit(lambda: ops.sdf_backward(view, 4, xq, gq, grid_grads, dec_grads[4], summed_scratch=scratch))
    summed = view.summed[4]
    v1 = ops.NetView([summed], [tuple(p.data for p in net.decoder_params(4))], math_mode=_lib.MATH_FP32)
    v1s = ops.NetView([summed], [tuple(p.data for p in net.decoder_params(4))], math_mode=_lib.MATH_FP32, summed=[summed])
    gg = [torch.zeros_like(summed, memory_format=torch.preserve_format)]
    sc = [torch.zeros_like(summed, memory_format=torch.preserve_format)]
    t_k = timeit(lambda: ops.sdf_backward(v1s, 0, xq, gq, gg, dec_grads[4], summed_scratch=sc))
    t_nog = timeit(lambda: ops.sdf_backward(v1s, 0, xq, gq, [None], dec_grads[4], summed_scratch=sc))
    loss = torch.zeros(1, device=dev)
    t_step = timeit(lambda: ops.sdf_train_step(view, 0x1f, xq, gq, 1.0 / n, grid_grads, dec_grads, loss, summed_scratch=scratch, scatter_scratch=ss))
    print(f"n={n:8d}: backward lod4 + cascade {t_full:.3f} ms | one-grid kernel + copy-out {t_k:.3f} | no grid grads {t_nog:.3f} | "
          f"5-head train step {t_step:.3f} ms", flush=True)
