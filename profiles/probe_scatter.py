"""L2 atomic (RED.v4.f32) scatter peak on this box (nglod_probe_scatter): the roof of the backward's grid-gradient scatter.
GB/s of reduced bytes for the backward's address stream (8 corner lines of a random cell, 8 lanes x red.v4 per line) over
grids of several sizes, as a function of the launch shape."""
import ctypes, sys, torch
sys.path.insert(0, '/root/repo')
from nglod_b200 import _lib
lib = _lib.load()
dev = torch.device('cuda', 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
def run(R, nq, smem, ctas, it=7, do_flush=True):
    buf = torch.zeros((R + 1) ** 3 * 32, device=dev)
    call = lambda seed: _lib.check(lib.nglod_probe_scatter(ctypes.c_void_p(buf.data_ptr()), R, nq, smem, ctas, seed, st()), "probe")
    for i in range(2): call(i)
    torch.cuda.synchronize()
    ts = []
    for i in range(it):
        if do_flush: flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); call(100 + i); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    t = ts[len(ts) // 2]
    return t, nq * 1024 / (t * 1e-3) / 1e9
print("R  n_queries  smem/CTA CTAs/SM   ms   GB/s reduced")
for R in (64, 32, 16, 8, 4):
    for nq in (1 << 20, 1 << 22):
        for smem, ctas in ((0, 0), (0, 1), (200 << 10, 1)):
            t, g = run(R, nq, smem, ctas)
            print(f"{R:3d} {nq:9d} {smem >> 10:5d} KB {ctas:3d}   {t:7.3f} {g:9.0f}", flush=True)
print("-- exactly k CTAs of 512 threads (R = 64, 2^20 queries): per-SM or chip-wide limit?")
for k in (148, 111, 74, 37, 18):
    t, g = run(64, 1 << 20, 0, -k)
    print(f"  {k:4d} CTAs  {t:7.3f} ms {g:9.0f} GB/s   {g / k:7.1f} GB/s per CTA", flush=True)
