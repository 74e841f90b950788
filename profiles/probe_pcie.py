"""Device -> pinned host copy bandwidth of the box (cudaMemcpyAsync, one stream), by size; and the same while a tracer frame runs."""
import sys, time, torch, numpy as np
sys.path.insert(0, '/root/repo')
dev = torch.device('cuda', 0)
for mb in (1, 4, 16, 52, 128):
    n = mb << 20
    src = torch.empty(n, dtype=torch.uint8, device=dev); dst = torch.empty(n, dtype=torch.uint8).pin_memory()
    for _ in range(3): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); dst.copy_(src, non_blocking=True); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    t = float(np.median(ts))
    print(f"D2H {mb:4d} MB: {t:.3f} ms = {n / t / 1e6:.1f} GB/s")
    src2 = torch.empty(n, dtype=torch.uint8).pin_memory(); dst2 = torch.empty(n, dtype=torch.uint8, device=dev)
    for _ in range(3): dst2.copy_(src2, non_blocking=True)
    torch.cuda.synchronize(); ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); dst2.copy_(src2, non_blocking=True); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    t = float(np.median(ts))
    print(f"H2D {mb:4d} MB: {t:.3f} ms = {n / t / 1e6:.1f} GB/s")
