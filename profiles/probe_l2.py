"""L2 -> SM gather peak on this box (nglod_probe_gather): the roofline denominator of the gather-bound SDF kernels.
Prints GB/s for the structured (8 corner lines of a random cell) and unstructured (random 128-byte lines) address
streams over a 35 MB (R=64) and a 4.6 MB (R=32) fp32 grid, with and without an L2 flush before the launch."""
import ctypes, sys, torch
sys.path.insert(0, '/root/repo')
from nglod_b200 import _lib
lib = _lib.load()
dev = torch.device('cuda', 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
sink = torch.zeros(4, dtype=torch.int32, device=dev)
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
def run(R, nq, infl, structured, do_flush, it=10):
    S = R + 1
    buf = torch.randn(S * S * S * 32, device=dev)
    call = lambda seed: _lib.check(lib.nglod_probe_gather(ctypes.c_void_p(buf.data_ptr()), R, nq, infl, structured, seed,
                                                          ctypes.c_void_p(sink.data_ptr()), st()), "probe")
    for i in range(3): call(i)
    torch.cuda.synchronize()
    ts = []
    for i in range(it):
        if do_flush: flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); call(100 + i); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return nq * 1024 / (ts[len(ts) // 2] * 1e-3) / 1e9, ts[0], ts[len(ts) // 2]
for R in (64, 32):
    for nq in (1 << 20, 1 << 23):
        for structured in (1, 0):
            for infl in (1, 2, 3):
                for fl in (True, False):
                    gbs, tmin, tmed = run(R, nq, infl, structured, fl)
                    print(f"R={R} nq=2^{nq.bit_length()-1} structured={structured} in_flight={infl} flush={int(fl)}: "
                          f"{gbs:9.1f} GB/s (median {tmed*1e3:.1f} us, min {tmin*1e3:.1f} us)", flush=True)
