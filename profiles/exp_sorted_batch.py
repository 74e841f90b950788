"""Does the order of a batch matter?  2^20-query backward / forward with the points in random order, sorted by the
cell of the 64^3 grid (row-major), and sorted along a Morton curve."""
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
def part1by2(v):
    v = v & 0x3ff
    v = (v | (v << 16)) & 0x30000ff
    v = (v | (v << 8)) & 0x300f00f
    v = (v | (v << 4)) & 0x30c30c3
    v = (v | (v << 2)) & 0x9249249
    return v
for n in (1 << 20, 500000):
    x = torch.rand(n, 3, device=dev, generator=g) * 2 - 1
    gq = torch.rand(n, device=dev, generator=g)
    cell = ((x + 1) * 32).floor().clamp(0, 63).long()
    rowmajor = (cell[:, 2] * 64 + cell[:, 1]) * 64 + cell[:, 0]
    morton = part1by2(cell[:, 0]) | (part1by2(cell[:, 1]) << 1) | (part1by2(cell[:, 2]) << 2)
    view = net.net_view(inference=False)
    grid_grads = [torch.zeros_like(f.fm.data, memory_format=torch.preserve_format) for f in net.features]
    scratch = net.summed_grad_scratch()
    dec_grads = [tuple(torch.zeros_like(p) for p in net.decoder_params(l)) for l in range(5)]
    loss = torch.zeros(1, device=dev)
    for tag, order in (("random", None), ("row-major cells", torch.argsort(rowmajor)), ("morton cells", torch.argsort(morton))):
        xs = x if order is None else x[order].contiguous()
        gs = gq if order is None else gq[order].contiguous()
        tb = timeit(lambda: ops.sdf_backward(view, 4, xs, gs, grid_grads, dec_grads[4], summed_scratch=scratch))
        tf = timeit(lambda: ops.sdf_forward(net.net_view(), 4, xs))
        tst = timeit(lambda: ops.sdf_train_step(view, 0x1f, xs, gs, 1.0 / n, grid_grads, dec_grads, loss, summed_scratch=scratch,
                                               scatter_scratch=net.scatter_scratch()))
        print(f"n={n:8d} {tag:16s}: backward lod4 {tb:.3f} ms | forward {tf * 1e3:.1f} us | 5-head step {tst:.3f} ms", flush=True)
    ts = timeit(lambda: torch.argsort(morton))
    print(f"   torch.argsort of the keys: {ts:.3f} ms")
