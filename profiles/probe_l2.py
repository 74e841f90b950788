"""L2 -> SM gather peak on this box (nglod_probe_gather): the roofline denominator of the gather-bound SDF kernels.
GB/s for the structured (8 corner lines of a random cell) address stream over the 35 MB lod-4 grid as a function of the
launch shape: shared-memory carve-out per CTA (what is left of the 228 KB is L1), CTAs per SM, loads in flight per lane."""
import ctypes, sys, torch
sys.path.insert(0, '/root/repo')
from nglod_b200 import _lib
lib = _lib.load()
dev = torch.device('cuda', 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
sink = torch.zeros(4, dtype=torch.int32, device=dev)
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
R = 64
buf = torch.randn((R + 1) ** 3 * 32, device=dev)
def run(nq, infl, structured, smem, ctas, it=7):
    call = lambda seed: _lib.check(lib.nglod_probe_gather(ctypes.c_void_p(buf.data_ptr()), R, nq, infl, structured, smem, ctas, seed,
                                                          ctypes.c_void_p(sink.data_ptr()), st()), "probe")
    for i in range(2): call(i)
    torch.cuda.synchronize()
    ts = []
    for i in range(it):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); call(100 + i); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return nq * 1024 / (ts[len(ts) // 2] * 1e-3) / 1e9
nq = 1 << 23
print("smem/CTA  CTAs/SM  threads/SM   GB/s at 8 / 16 / 24 loads in flight per lane (structured), 8 (random lines)")
for smem, ctas in ((0, 0), (0, 2), (0, 1), (56 << 10, 0), (112 << 10, 2), (112 << 10, 1), (160 << 10, 1), (192 << 10, 1), (208 << 10, 1),
                   (216 << 10, 1), (226 << 10, 1)):
    vals = [run(nq, f, 1, smem, ctas) for f in (1, 2, 3)] + [run(nq, 1, 0, smem, ctas)]
    per_sm = ctas if ctas else (228 * 1024 // max(smem + 1024, 1) if smem else 4)
    print(f"{smem >> 10:5d} KB  {ctas:5d}  {min(per_sm, 4) * 512:8d}   " + " / ".join(f"{v:8.0f}" for v in vals), flush=True)
