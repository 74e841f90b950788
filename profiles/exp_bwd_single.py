import sys, torch, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from nglod_b200 import ops, _lib
from helpers import rand5_model
dev = torch.device('cuda', 0)
net, args = rand5_model(dev)
g = torch.Generator(device=dev).manual_seed(1)
n = 1 << 20
xq = torch.rand(n, 3, device=dev, generator=g) * 2 - 1
gq = torch.rand(n, device=dev, generator=g)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, it=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(it):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
view = net.net_view(inference=False)
grid_grads = [torch.zeros_like(f.fm.data, memory_format=torch.preserve_format) for f in net.features]
dec_grad = tuple(torch.zeros_like(p) for p in net.decoder_params(4))
print("multi-LOD backward", timeit(lambda: ops.sdf_backward(view, 4, xq, gq, grid_grads, dec_grad)))
summed = net.net_view().summed[4]
v1 = ops.NetView([summed], [tuple(p.data for p in net.decoder_params(4))], math_mode=_lib.MATH_FP32)
gg = [torch.zeros_like(summed, memory_format=torch.preserve_format)]
print("single-grid backward", timeit(lambda: ops.sdf_backward(v1, 0, xq, gq, gg, dec_grad)))
print("single-grid backward, no grid grads", timeit(lambda: ops.sdf_backward(v1, 0, xq, gq, [None], dec_grad)))
print("multi backward, no grid grads", timeit(lambda: ops.sdf_backward(view, 4, xq, gq, [None]*5, dec_grad)))
