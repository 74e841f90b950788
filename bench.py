#!/usr/bin/env python
"""bench.py -- headline benchmark of the NGLOD hot path on N B200s (one process per GPU).

  python bench.py --gpus 1 --steps 20 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...      # the UNMODIFIED reference's CPU path on the box's host cores

Workload (BASELINE.json configs[1]): SphereTracer.forward over a 1280x720 perspective frame at lod 4 of a 5-LOD
OctreeSDF (feature-dim 32, hidden 128) fitted IN-RUN to a torus.  Both arms trace the SAME net and the SAME rays: the fit
is a seeded plain-torch Adam fit (no kernel of this repo, nothing from oracle/) whose result is cached in
baseline/_fit_cache.pt by whichever arm runs first on the box; the rays come from look_at on the host under one seed.
A "step" is one full frame (921 600 rays) per GPU; with N GPUs every rank traces its own frame (a different camera
azimuth), no data-path collective -> "scaling": "weak"; value = N * 921600 / max-over-ranks time.  The partitionings
north_star names (one 4K frame in screen strips with the final gather; the data-parallel training step with its gradient
exchange) are strong-scaling measurements and ride along under "strong_scaling" at every N.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

W, H = 1280, 720
LOD = 4
NUM_LODS = 5
CAM_FROM, CAM_TO, FOV = [-2.8, 2.8, -2.8], [0.0, 0.0, 0.0], 30.0
FIT_STEPS, FIT_BATCH, FIT_SEED = 300, 65536, 7
SDF_N = 1 << 20
MATH_MODE = os.environ.get("NGLOD_MATH", "tc")      # "tc" = tcgen05 3xTF32 decoder, "fp32" = CUDA cores
GRID_STORAGE = os.environ.get("NGLOD_GRID_STORAGE", "fp32")   # "fp32" (headline) | "fp16" x-pair lines (extras)
SUM_LODS = os.environ.get("NGLOD_SUM_LODS", "1") != "0"        # inference gathers the prefix-summed grid of the LOD
# Algorithmic gather bytes per SDF evaluation.  SURVEY.md 8d quotes the per-LOD formulation ((lod+1) x 8 corners x 32 ch
# x 4 B = 5120 B); the inference kernels evaluate the same function from ONE prefix-summed grid,
# so the bytes the algorithm has to move are 8 corner lines x 128 B = 1024 B (fp32) per evaluation.
GATHER_BYTES_PER_QUERY_PER_LOD = (LOD + 1) * 8 * 32 * 4
GATHER_BYTES_PER_QUERY = 8 * 32 * 4 if SUM_LODS else GATHER_BYTES_PER_QUERY_PER_LOD
IO_BYTES_PER_QUERY = 16
RAY_IO_BYTES = 24 + 12 + 4 + 1 + 12                        # ray_o, ray_d in; x, depth, hit, normal out
TENSOR_FLOP_PER_EVAL = 2 * 32 * 128 * 3                    # the tensor-eligible 32x128 contraction, 3 TF32 passes
WORKLOAD = ("SphereTracer.forward 1280x720 persp fov30 lod4, OctreeSDF num-lods=5 feature-dim=32 hidden=128 fitted in-run "
            "to a torus (BASELINE.json configs[1])")


def fit_plan():
    """(device, steps, batch, cache path) of the shared fit; the cache name carries the plan, so a fit made under another
    plan (a CPU-only box, a test's shortened set-up) is never mistaken for this one."""
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    steps, batch = (FIT_STEPS, FIT_BATCH) if dev == "cuda" else (60, 8192)
    steps = int(os.environ.get("NGLOD_FIT_STEPS", steps))               # tests shrink the (untimed) set-up
    return dev, steps, batch, os.path.join(ROOT, "baseline", f"_fit_cache_{dev}_s{steps}_b{batch}_seed{FIT_SEED}.pt")


def shared_config(world):
    """The `config` both arms print (identical by construction: same workload, same net, same rays)."""
    return {"workload": WORKLOAD, "rays_per_gpu_step": W * H, "num_steps": 256, "lod": LOD,
            "fit": "seeded plain-torch Adam fit ({} steps x {} points), analytic torus SDF labels, shared between the arms "
                   "through baseline/_fit_cache*.pt".format(*fit_plan()[1:3]),
            "rays": "look_at([-2.8,2.8,-2.8] rotated about y by 360*rank/N, [0,0,0], 1280, 720, fov 30) on the host, seed 1000+rank",
            "l2": "flushed between timed iterations (256 MB memset)", "parallelism": f"one frame per GPU x{world}, no collective"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), float(j.get("bf16_tflops_sustained", 1400.9)), "measured"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, 1400.0, "fallback"


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, from the committed `ncu --set full` summary
    (profiles/ncu_traffic.json, written by profiles/ncu_traffic.py from the .ncu-rep of the same build); None if absent."""
    try:
        j = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        e = j[kernel]
        return int(e["dram_bytes_read"]) + int(e["dram_bytes_write"])
    except Exception:  # noqa: BLE001
        return None


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock / throttle reasons during the timed region (NVML; falls back to one nvidia-smi poll)."""

    def __init__(self, index):
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._th = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.005)

    def start(self):
        if self.nv is not None:
            self._th = threading.Thread(target=self._loop, daemon=True)
            self._th.start()

    def stop(self):
        self._stop.set()
        if self._th is not None:
            self._th.join()
        if self.nv is None:
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                    capture_output=True, text=True, timeout=10).stdout.strip().split(",")
                self.samples, self.max_mhz = [int(out[0])], int(out[1])
            except Exception:  # noqa: BLE001
                pass
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------- shared set-up (both arms)
def torch_reference_sdf_all(sd, x, num_lods):
    """Plain-torch statement of OctreeSDF.sdf(x, return_lst=True) on a state_dict (F.grid_sample per LOD, running sum,
    cat [x, feat], Linear-ReLU-Linear of each head; sdf-net/lib/models/OctreeSDF.py:46-57,94-153).  Used ONLY for the untimed
    in-run fit, by both arms -- it calls no kernel of this repository and imports nothing from oracle/."""
    import torch.nn.functional as F
    grid = x.reshape(1, -1, 1, 1, 3)
    feat = 0
    outs = []
    for l in range(num_lods):
        s = F.grid_sample(sd[f"features.{l}.fm"], grid, align_corners=True, padding_mode="border")[0, :, :, 0, 0].t()
        feat = feat + s
        h = torch.relu(F.linear(torch.cat([x, feat], dim=-1), sd[f"louts.{l}.0.weight"], sd[f"louts.{l}.0.bias"]))
        outs.append(F.linear(h, sd[f"louts.{l}.2.weight"], sd[f"louts.{l}.2.bias"]))
    return outs


def fit_shared(init_state_dict, log):
    """The in-run fit both arms trace: seeded Adam (lr 1e-3) on all LOD heads (trainer.py:317-339's loss) against the
    analytic torus SDF (R .6, r .25), half of every batch pulled near the surface like MeshDataset's near / trace modes.
    Cached on disk so that the second arm on the same box traces bit-identical weights."""
    dev, steps, batch, FIT_CACHE = fit_plan()
    if os.path.exists(FIT_CACHE):
        try:
            sd = torch.load(FIT_CACHE, map_location="cpu")
            if set(sd) == set(init_state_dict) and all(sd[k].shape == init_state_dict[k].shape for k in sd):
                log(f"fit: loaded {FIT_CACHE} (fitted by the arm that ran first on this box)")
                return sd, "cache"
        except Exception:  # noqa: BLE001
            pass
    sd = {k: v.detach().clone().to(dev).contiguous().requires_grad_(True) for k, v in init_state_dict.items()}
    opt = torch.optim.Adam(list(sd.values()), lr=1e-3)
    g = torch.Generator(device=dev).manual_seed(FIT_SEED)
    t0 = time.time()
    last = 0.0
    for _ in range(steps):
        p = torch.rand(batch, 3, device=dev, generator=g) * 2 - 1
        surf = p[: batch // 2]
        q = torch.sqrt(surf[:, 0] ** 2 + surf[:, 2] ** 2)
        ring = torch.stack([surf[:, 0] / q * 0.6, torch.zeros_like(q), surf[:, 2] / q * 0.6], dim=1)
        dirv = torch.nn.functional.normalize(surf - ring, dim=1)
        p = torch.cat([ring + dirv * (0.25 + 0.01 * torch.randn(batch // 2, 1, device=dev, generator=g)), p[batch // 2:]])
        qq = torch.sqrt(p[:, 0] ** 2 + p[:, 2] ** 2) - 0.6
        gt = (torch.sqrt(qq * qq + p[:, 1] ** 2) - 0.25).unsqueeze(1)
        opt.zero_grad(set_to_none=True)
        loss = sum(((d - gt) ** 2).sum() for d in torch_reference_sdf_all(sd, p, NUM_LODS)) / batch
        loss.backward()
        opt.step()
        last = loss.detach()
    out = {k: v.detach().cpu().contiguous() for k, v in sd.items()}
    log(f"fit: {steps} Adam steps x {batch} pts on {dev} in {time.time() - t0:.1f}s, final loss {float(last):.3e}")
    try:
        os.makedirs(os.path.dirname(FIT_CACHE), exist_ok=True)
        tmp = FIT_CACHE + f".{os.getpid()}"
        torch.save(out, tmp)
        os.replace(tmp, FIT_CACHE)
    except Exception:  # noqa: BLE001
        pass
    return out, "fitted"


def camera_from(azimuth_deg):
    a = math.radians(azimuth_deg)
    return [CAM_FROM[0] * math.cos(a) + CAM_FROM[2] * math.sin(a), CAM_FROM[1],
            -CAM_FROM[0] * math.sin(a) + CAM_FROM[2] * math.cos(a)]


def make_rays(device, azimuth_deg=0.0, seed=0):
    """look_at rays for the 720p frame, generated ON THE HOST under torch.manual_seed(1000 + seed) so that both arms (and
    the reference's own look_at, which draws the same torch.rand sequence) get the same rays; moved to `device`."""
    from nglod_b200.lib.geoutils import look_at
    torch.manual_seed(1000 + seed)
    o, d = look_at(camera_from(azimuth_deg), CAM_TO, W, H, mode="persp", fov=FOV, device="cpu")
    return o.to(device).contiguous(), d.to(device).contiguous()


def build_and_fit(device, log):
    """5-LOD OctreeSDF carrying the shared in-run fit (fit_shared), on `device`."""
    from nglod_b200.lib.options import parse_options
    from nglod_b200.lib.models import OctreeSDF
    args = parse_options(return_parser=True).parse_args(
        ["--net", "OctreeSDF", "--num-lods", str(NUM_LODS), "--feature-dim", "32", "--lod", str(LOD),
         "--render-res", str(W), str(H)])
    torch.manual_seed(0)
    net = OctreeSDF(args)
    sd, how = fit_shared({k: v.detach().clone().contiguous() for k, v in net.state_dict().items()}, log)
    net.load_state_dict(sd)
    net = net.to(device)
    net.math_mode = MATH_MODE
    net.grid_storage, net.sum_lods = GRID_STORAGE, SUM_LODS
    net.lod = LOD
    net.eval()
    net.fit_source = how
    return net, args


def flush_l2(buf):
    buf.zero_()


def measure_l2_gather_peak(device):
    """The roofline denominator of the gather-bound kernels, measured on this box in this run (nglod_probe_gather): the
    address stream of the SDF gather (8 corner lines of a pseudo-random cell of a 35 MB channels-last grid, 8 lanes x
    LDG.128 per line), no arithmetic.  `best` = the launch shape that delivers most (no shared-memory carve-out, full
    occupancy); `kernel_shape` = one 512-thread CTA per SM next to a 225 KB carve-out, i.e. what is left to a kernel that
    keeps its tcgen05 operand ring in shared memory."""
    import ctypes
    from nglod_b200 import _lib
    lib = _lib.load()
    R = 64
    buf = torch.randn((R + 1) ** 3 * 32, device=device)
    sink = torch.zeros(4, dtype=torch.int32, device=device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    nq = 1 << 22

    def run(infl, smem, ctas):
        def call(seed):
            _lib.check(lib.nglod_probe_gather(ctypes.c_void_p(buf.data_ptr()), R, nq, infl, 1, smem, ctas, seed,
                                              ctypes.c_void_p(sink.data_ptr()),
                                              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "nglod_probe_gather")
        for i in range(2):
            call(i)
        torch.cuda.synchronize()
        ts = []
        for i in range(5):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            call(100 + i)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return nq * 1024 / (float(np.median(ts)) * 1e-3) / 1e9
    with torch.cuda.device(device):
        best = max(run(2, 0, 0), run(3, 0, 0))
        shape = max(run(2, 225 << 10, 1), run(1, 225 << 10, 1))
    return {"best_GBs": best, "kernel_shape_GBs": shape,
            "how": "nglod_probe_gather, 2^22 cells x 8 lines x 128 B from a 35 MB grid, L2 flushed before each launch, median of 5"}


def measure_scatter_peak(device):
    """The roofline denominator of the backward's grid-gradient scatter, measured on this box in this run
    (nglod_probe_scatter): the backward's address stream (8 corner lines of a pseudo-random cell of the 35 MB grid, 8 lanes x
    red.global.add.v4.f32 per line), no arithmetic.  What the L2 atomic units take does not depend on the launch shape."""
    import ctypes
    from nglod_b200 import _lib
    lib = _lib.load()
    R = 64
    buf = torch.zeros((R + 1) ** 3 * 32, device=device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    nq = 1 << 22
    ts = []
    with torch.cuda.device(device):
        for i in range(7):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.check(lib.nglod_probe_scatter(ctypes.c_void_p(buf.data_ptr()), R, nq, 0, 0, 100 + i,
                                               ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "nglod_probe_scatter")
            b.record()
            torch.cuda.synchronize()
            if i >= 2:
                ts.append(a.elapsed_time(b))
    return {"GBs": nq * 1024 / (float(np.median(ts)) * 1e-3) / 1e9,
            "how": "nglod_probe_scatter, 2^22 cells x 8 lines x 128 B of red.global.add.v4.f32 into a 35 MB grid, L2 flushed "
                   "before each launch, median of 5"}


# ----------------------------------------------------------------------------------------------- ours
def run_ours(ns):
    from nglod_b200 import dist as ndist
    from nglod_b200 import ops
    from nglod_b200.lib.tracer import SphereTracer
    rank, world, local = ndist.init_from_env()
    numa = ndist.bind_to_gpu_numa_node(local)        # before any pinned allocation (first touch)
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)

    def log(msg):
        if rank == 0:
            print("[bench] " + msg, file=sys.stderr, flush=True)

    if world > 1 and rank != 0:                       # one rank fits (or loads the cache), the others take its weights
        torch.distributed.barrier()
    net, args = build_and_fit(device, log)
    if world > 1 and rank == 0:
        torch.distributed.barrier()
    if world > 1:                       # identical weights on every rank, whatever the cache state
        for p in net.parameters():
            torch.distributed.broadcast(p.data, src=0)
        net.mark_grids_dirty()
    tracer = SphereTracer(args)
    azimuth = 360.0 * rank / max(world, 1)
    ray_o, ray_d = make_rays(device, azimuth_deg=azimuth, seed=rank)
    n_rays = ray_o.shape[0]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)      # 256 MB > 126 MB L2
    view = net.net_view()

    # one instrumented call: SDF evaluations per frame (for the algorithmic byte count)
    stats = torch.zeros(2, dtype=torch.int64, device=device)
    x, depth, hit, normal = ops.sphere_trace(view, LOD, ray_o, ray_d, stats=stats)
    torch.cuda.synchronize()
    n_eval, n_march = int(stats[0]), int(stats[1])
    n_hit = int(hit.sum())
    log(f"frame: {n_rays} rays, {n_hit} hits, {n_eval} sdf evals ({n_eval / n_rays:.2f}/ray), {n_march} march steps")

    def step():
        return tracer(net, ray_o, ray_d)

    for _ in range(max(ns.warmup, 3)):
        flush_l2(flush)
        step()
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    clocks.start()
    evs = []
    for _ in range(ns.steps):
        flush_l2(flush)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = ndist.max_over_ranks(sum(step_ms), device)
    kernel_ms = float(np.mean(step_ms))          # the step IS one sphere_trace_kernel launch (+ a 4-byte memset)

    # ---- e2e, the call a user makes (Renderer.render_lookat + .cpu() in the reference): camera + jittered window in
    #      pinned host memory -> rays generated on the device -> trace -> depth / hit / normal into pinned host memory
    from nglod_b200.lib.geoutils import _window
    torch.manual_seed(1000 + rank)
    wx, wy = _window(W, H, "cpu")
    wx, wy = wx.pin_memory(), wy.pin_memory()
    cam_fields = ("depth", "hit", "normal")
    cam_out = {k: torch.empty(s, dtype=dt).pin_memory() for k, s, dt in
               (("depth", (n_rays, 1), torch.float32), ("hit", (n_rays,), torch.bool), ("normal", (n_rays, 3), torch.float32))}
    cam_f = camera_from(azimuth)

    cam_packed = {}

    def e2e_step():         # ONE library call: the tracer writes 16-byte {depth, normal} records + hit bytes into pinned host memory
        tracer.trace_lookat_host(net, cam_f, CAM_TO, W, H, fov=FOV, mode="persp", window=(wx, wy), out=cam_packed,
                                 fields=cam_fields, packed=True)

    def e2e_chunked_step():  # round 2's first path: 3 chunks on 2 streams + cudaMemcpyAsync of the fields
        tracer.trace_lookat_host(net, cam_f, CAM_TO, W, H, fov=FOV, mode="persp", window=(wx, wy), out=cam_out, fields=cam_fields)

    def wall(fn, iters):
        for _ in range(3):
            fn()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            fn()
        return ndist.max_over_ranks(time.perf_counter() - t0, device)
    e2e_s = wall(e2e_step, ns.steps)
    cam_hits = int(cam_packed["hit"].sum())
    e2e_chunked_s = wall(e2e_chunked_step, ns.steps)
    if not (torch.equal(cam_packed["hit"], cam_out["hit"]) and torch.equal(cam_packed["packed"][:, 0:1], cam_out["depth"])
            and torch.equal(cam_packed["packed"][:, 1:4], cam_out["normal"])):
        raise RuntimeError("bench: the packed and the chunked host paths disagree")

    # ---- the same frame with HOST RAY BUFFERS in and every RenderBuffer field out (round 1's e2e leg): 22 MB up, 27 MB down
    ho, hd = ray_o.cpu().pin_memory(), ray_d.cpu().pin_memory()
    out_host = {k: torch.empty(s, dtype=dt).pin_memory() for k, s, dt in
                (("x", (n_rays, 3), torch.float32), ("depth", (n_rays, 1), torch.float32),
                 ("hit", (n_rays,), torch.bool), ("normal", (n_rays, 3), torch.float32))}
    e2e_rays_s = wall(lambda: tracer.trace_host(net, ho, hd, out=out_host), ns.steps)
    clock_info = clocks.stop()

    # ---- SDF query throughput (BASELINE.json configs[0]): 2^20 random points, lod 4, forward and forward+backward
    g = torch.Generator(device=device).manual_seed(1)
    xq = torch.rand(SDF_N, 3, device=device, generator=g) * 2 - 1
    gq = torch.rand(SDF_N, device=device, generator=g)

    def time_kernel(fn, iters, do_flush=True):
        for _ in range(3):
            if do_flush:
                flush_l2(flush)
            fn()
        torch.cuda.synchronize()
        ev = []
        for _ in range(iters):
            if do_flush:
                flush_l2(flush)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            ev.append((a, b))
        torch.cuda.synchronize()
        return float(np.mean([a.elapsed_time(b) for a, b in ev]))

    fwd_ms = time_kernel(lambda: ops.sdf_forward(view, LOD, xq), max(ns.steps, 10))
    grid_grads = [torch.zeros_like(f.fm.data, memory_format=torch.preserve_format) for f in net.features]
    dec_grad = tuple(torch.zeros_like(p) for p in net.decoder_params(LOD))
    scratch = net.summed_grad_scratch() if view.summed is not None else None     # single-grid backward
    bwd_ms = time_kernel(lambda: ops.sdf_backward(view, LOD, xq, gq, grid_grads, dec_grad, summed_scratch=scratch),
                         max(ns.steps // 2, 5))
    fwd_ms = ndist.max_over_ranks(fwd_ms, device)
    bwd_ms = ndist.max_over_ranks(bwd_ms, device)
    # inputs larger than L2 instead of a flush: 16 distinct 2^20-point batches (201 MB of coordinates) in rotation --
    # the 35 MB summed grid then stays L2-resident between launches, which is the steady state of a query server
    xs_rot = [torch.rand(SDF_N, 3, device=device, generator=g) * 2 - 1 for _ in range(16)]
    rot = {"i": 0}

    def fwd_rot():
        ops.sdf_forward(view, LOD, xs_rot[rot["i"] % 16])
        rot["i"] += 1
    fwd_rot_ms = ndist.max_over_ranks(time_kernel(fwd_rot, 32, do_flush=False), device)
    del xs_rot

    l2peak = measure_l2_gather_peak(device) if rank == 0 else None
    scatter_peak = measure_scatter_peak(device) if rank == 0 else None

    variants = {}
    if not ns.no_extras:
        for tag, storage, summ in (("fp16_xpair_lines", "fp16", True), ("per_lod_gather", "fp32", False)):
            net.grid_storage, net.sum_lods = storage, summ
            v = net.net_view()
            f_ms = ndist.max_over_ranks(time_kernel(lambda: ops.sdf_forward(v, LOD, xq), 10), device)
            t_ms = ndist.max_over_ranks(time_kernel(lambda: tracer(net, ray_o, ray_d), 10), device)
            variants[tag] = {"forward_qps": world * SDF_N / (f_ms / 1e3), "forward_ms": f_ms,
                             "trace_rays_per_s": world * n_rays / (t_ms / 1e3), "trace_ms": t_ms}
        net.grid_storage, net.sum_lods = GRID_STORAGE, SUM_LODS
        # a renderer that keeps TWO frames in flight (alternating streams): the next frame's rays fill the SMs the
        # previous frame's straggler rays leave idle.  Throughput only -- the headline `value` is one frame at a time.
        s2 = [torch.cuda.Stream(device), torch.cuda.Stream(device)]
        outs = [tuple(torch.empty_like(t) for t in (x, depth, hit, normal)) for _ in range(2)]
        qs = torch.zeros(2, dtype=torch.int32, device=device)

        def frames_in_flight(k):
            for sidx in range(2):
                s2[sidx].wait_stream(torch.cuda.current_stream(device))
            for f in range(k):
                with torch.cuda.stream(s2[f % 2]):
                    ops.sphere_trace(view, LOD, ray_o, ray_d, out=outs[f % 2], queue=qs[f % 2:f % 2 + 1])
            for sidx in range(2):
                torch.cuda.current_stream(device).wait_stream(s2[sidx])
        frames_in_flight(4)
        torch.cuda.synchronize()
        flush_l2(flush)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        frames_in_flight(20)
        b.record()
        torch.cuda.synchronize()
        pipe_ms = ndist.max_over_ranks(a.elapsed_time(b) / 20, device)
        variants["two_frames_in_flight"] = {"trace_rays_per_s": world * n_rays / (pipe_ms / 1e3), "ms_per_frame": pipe_ms,
                                            "note": "20 frames on two alternating streams, one L2 flush before the batch"}
        xbig = torch.rand(1 << 23, 3, device=device, generator=g) * 2 - 1
        big_ms = ndist.max_over_ranks(time_kernel(lambda: ops.sdf_forward(view, LOD, xbig), 5), device)
        variants["forward_2^23_queries"] = {"forward_qps": world * (1 << 23) / (big_ms / 1e3), "forward_ms": big_ms}
        del xbig
    strong = run_strong_scaling(net, args, device, rank, world, flush, log)
    extras = {} if ns.no_extras else run_extras(net, args, device, rank, world, flush, log)
    if variants:
        extras["inference_variants"] = variants
    if rank == 0 and world == 1 and not ns.no_extras and not ns.no_cpu_baseline:
        extras["reference_on_this_gpu"] = reference_on_this_gpu(net, log)
    if rank != 0:
        return
    hbm_peak, tensor_peak, peak_src = load_peaks()
    value = world * n_rays * ns.steps / (total_ms / 1e3)
    ks = kernel_ms / 1e3
    gather_GBs = n_eval * GATHER_BYTES_PER_QUERY / ks / 1e9                  # L2 -> SM bytes the gather has to move
    alg_bytes = n_eval * (GATHER_BYTES_PER_QUERY + IO_BYTES_PER_QUERY) + n_rays * RAY_IO_BYTES
    traffic = ncu_traffic("sphere_trace_kernel")
    q_gather_GBs = SDF_N * GATHER_BYTES_PER_QUERY / (fwd_ms / 1e3) / 1e9
    fwd_kernel = "sdf_forward_ws_kernel" if MATH_MODE == "tc" else "sdf_forward_kernel"
    fwd_traffic = ncu_traffic(fwd_kernel)
    line = {
        "metric": "sphere_traced_rays_per_sec_1280x720_lod4",
        "value": value,
        "unit": "rays/s",
        "n_gpus": world,
        "steps": ns.steps,
        "warmup": max(ns.warmup, 3),
        "ms_per_step": total_ms / ns.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": shared_config(world),
        "sdf_queries_per_sec": world * SDF_N / (fwd_ms / 1e3),          # the other half of BASELINE.json's metric
        "sdf_forward_backward_queries_per_sec": world * SDF_N / ((fwd_ms + bwd_ms) / 1e3),
        "fps_per_gpu": 1e3 / (total_ms / ns.steps),
        "details": {"sdf_evals_per_ray": n_eval / n_rays, "hit_fraction": n_hit / n_rays, "math_mode": MATH_MODE,
                    "grid_storage": GRID_STORAGE, "fit_source": net.fit_source, "numa_node": numa,
                    "lod_sum": "prefix-summed grid (one 8-corner gather per evaluation)" if SUM_LODS else "per-LOD gather"},
        "e2e": {"value": world * n_rays * ns.steps / e2e_s, "unit": "rays/s",
                "h2d_bytes_per_step": (W + H) * 4 + 64, "d2h_bytes_per_step": n_rays * (4 + 1 + 12),
                "ms_per_step": e2e_s / ns.steps * 1e3, "hits": cam_hits,
                "api": "SphereTracer.trace_lookat_host(packed=True): camera pose + jittered window (pinned host) in, depth / hit / "
                       "normal (pinned host) out; rays are generated on the device, as the reference's Renderer.render_lookat "
                       "does; ONE library call (nglod_sphere_trace_camera) whose tracer kernel writes each ray's 16-byte "
                       "{depth, normal} record + hit byte into the pinned host buffers as the ray retires (posted PCIe "
                       "writes: the d2h bytes cross the bus inside the kernel, no copy afterwards)"},
        "e2e_chunked": {"value": world * n_rays * ns.steps / e2e_chunked_s, "unit": "rays/s",
                        "h2d_bytes_per_step": (W + H) * 4 + 64, "d2h_bytes_per_step": n_rays * (4 + 1 + 12),
                        "ms_per_step": e2e_chunked_s / ns.steps * 1e3,
                        "api": "SphereTracer.trace_lookat_host(packed=False): 3 ray ranges on 2 streams, fields copied with "
                               "cudaMemcpyAsync as each range finishes (identical results, checked in-run)"},
        "e2e_ray_buffers": {"value": world * n_rays * ns.steps / e2e_rays_s, "unit": "rays/s",
                            "h2d_bytes_per_step": n_rays * 24, "d2h_bytes_per_step": n_rays * (12 + 4 + 1 + 12),
                            "ms_per_step": e2e_rays_s / ns.steps * 1e3,
                            "api": "SphereTracer.trace_host: ray_o / ray_d from pinned host memory, x / depth / hit / normal back"},
        "gpu_launches": ns.steps,
        "clocks": clock_info,
        "roofline": {"bound": "l2", "kernel": "sphere_trace_kernel", "achieved": gather_GBs, "peak": l2peak["best_GBs"],
                     "unit": "GB/s", "frac": gather_GBs / l2peak["best_GBs"], "traffic": traffic,
                     "peak_source": "measured in this run: " + l2peak["how"],
                     "peak_at_kernel_launch_shape_GBs": l2peak["kernel_shape_GBs"],
                     "frac_of_kernel_shape_peak": gather_GBs / l2peak["kernel_shape_GBs"],
                     "kernel_ms": kernel_ms, "sdf_evals_per_launch": n_eval,
                     "tensor": {"achieved_TFLOPs": n_eval * TENSOR_FLOP_PER_EVAL / ks / 1e12, "peak_TFLOPs": tensor_peak,
                                "frac": n_eval * TENSOR_FLOP_PER_EVAL / ks / 1e12 / tensor_peak,
                                "note": "SURVEY 8d form: evals x 2*32*128 x 3 TF32 passes over the measured bf16 GEMM peak"},
                     "dram": {"traffic_GBs": (traffic / ks / 1e9) if traffic else None, "peak_GBs": hbm_peak,
                              "frac": (traffic / ks / 1e9 / hbm_peak) if traffic else None, "peak_source": peak_src},
                     "hbm_form_GBs": alg_bytes / ks / 1e9,
                     "per_lod_formulation_GBs": (n_eval * (GATHER_BYTES_PER_QUERY_PER_LOD + IO_BYTES_PER_QUERY)
                                                 + n_rays * RAY_IO_BYTES) / ks / 1e9,
                     "note": f"achieved = sdf_evals x {GATHER_BYTES_PER_QUERY} B (the 8 corner lines of the prefix-summed grid, DESIGN "
                             "4.0) / kernel time: the gather is served by L2 (the 35 MB grid is resident; DRAM traffic is the "
                             "grid once + ray I/O), so the roof is the L2 -> SM path measured by the probe, not HBM. "
                             "hbm_form_GBs / per_lod_formulation_GBs keep round 1's and SURVEY 8d's byte counts for "
                             "comparison. The frame is bound by the latency of its longest rays (256 dependent steps), "
                             "not by throughput: DESIGN.md section 4"},
        "sdf_queries": {"n": SDF_N, "lod": LOD,
                        "forward_qps": world * SDF_N / (fwd_ms / 1e3), "forward_ms": fwd_ms,
                        "forward_backward_qps": world * SDF_N / ((fwd_ms + bwd_ms) / 1e3), "backward_ms": bwd_ms,
                        "forward_qps_inputs_rotating": world * SDF_N / (fwd_rot_ms / 1e3), "forward_ms_inputs_rotating": fwd_rot_ms,
                        "l2": "forward_ms: L2 flushed before every launch; *_inputs_rotating: no flush, 16 distinct input "
                              "batches (201 MB) in rotation, the summed grid stays L2-resident",
                        "roofline": {"bound": "l2", "kernel": fwd_kernel, "achieved": q_gather_GBs, "peak": l2peak["best_GBs"],
                                     "unit": "GB/s", "frac": q_gather_GBs / l2peak["best_GBs"],
                                     "frac_of_kernel_shape_peak": q_gather_GBs / l2peak["kernel_shape_GBs"],
                                     "traffic": fwd_traffic,
                                     "tensor_frac": SDF_N * TENSOR_FLOP_PER_EVAL / (fwd_ms / 1e3) / 1e12 / tensor_peak},
                        "backward_roofline": {
                            "bound": "l2_atomics", "kernel": "sdf_backward_tc_kernel",
                            "achieved": SDF_N * GATHER_BYTES_PER_QUERY / (bwd_ms / 1e3) / 1e9, "peak": scatter_peak["GBs"],
                            "unit": "GB/s", "frac": SDF_N * GATHER_BYTES_PER_QUERY / (bwd_ms / 1e3) / 1e9 / scatter_peak["GBs"],
                            "traffic": ncu_traffic("sdf_backward_tc_kernel"),
                            "peak_source": "measured in this run: " + scatter_peak["how"],
                            "note": "achieved = queries x 1024 B reduced into the gradient of the prefix-summed grid (8 corner "
                                    "lines x 128 B of red.global.add.v4.f32 per query) / backward_ms, which also contains the "
                                    "restriction cascade (~0.07 ms); the same 1024 B per query are gathered for the forward "
                                    "recompute on top of that"}},
        "strong_scaling": strong,
    }
    line["extras"] = extras
    if world == 1 and not ns.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_trace(net, ray_o.cpu(), ray_d.cpu(), budget_s=20.0, log=log)
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- strong scaling (configs 3, 5)
def run_strong_scaling(net, args, device, rank, world, flush, log):
    """The two partitionings north_star names, at this N, total work fixed (divide the N=1 time by this one for the
    speed-up): (a) ONE 3840x2160 frame with shadows in interleaved screen strips, final NCCL gather to rank 0 inside the
    timed region; (b) the data-parallel training step on a 500 000-point batch: per-rank sampling + mesh2sdf labels
    (prefetched on a second stream), fused 5-head step, reduce-scatter / Adam on 1/N / all-gather."""
    import copy
    from nglod_b200 import dist as ndist
    from nglod_b200 import ops
    from nglod_b200.lib.renderer import Renderer
    from nglod_b200.lib.tracer import SphereTracer
    from nglod_b200.lib.trainer import FusedTrainer
    from nglod_b200.lib.torchgp import torus, point_sample, normalize
    out = {"n_gpus": world}

    def timed(fn, iters=5, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        ts = []
        for _ in range(iters):
            flush_l2(flush)
            torch.cuda.synchronize()
            if world > 1:
                torch.distributed.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return ndist.max_over_ranks(float(np.median(ts)), device)

    # ---- (a) config 5: 3840x2160, shadows + normals; 4 interleaved strips per rank (silhouette columns cost more than empty
    #      ones); every rank renders its strips as one batch; depth / hit / normal / shadow (18 B per ray) gathered to rank 0
    w4, h4 = 3840, 2160
    torch.manual_seed(5)
    from nglod_b200.lib.geoutils import look_at
    ro, rd = look_at(CAM_FROM, CAM_TO, w4, h4, mode="persp", fov=FOV, device=device)
    strips_all = [ndist.interleaved_strips(w4, r, world, strips_per_rank=4) for r in range(world)]
    mine = strips_all[rank]
    ro = torch.cat([ro[c0 * h4:c1 * h4] for c0, c1 in mine]).contiguous()
    rd = torch.cat([rd[c0 * h4:c1 * h4] for c0, c1 in mine]).contiguous()
    rargs = copy.copy(args)
    rargs.shadow, rargs.ground_height, rargs.render_res = True, -0.4, [ro.shape[0] // h4, h4]
    renderer = Renderer(SphereTracer(rargs), args=rargs, device=device)
    frame = {}

    def render_and_gather():
        rb = renderer.render(net, ro, rd)
        # the final gather, field by field (no packing pass): depth 4 B, normal 12 B, hit 1 B, shadow 1 B per ray
        frame["full"] = [ndist.gather_strips(t, strips_all, h4, rank, world, dst=0) for t in
                         (rb.depth.reshape(-1, 1), rb.normal.reshape(-1, 3), rb.hit.reshape(-1, 1).view(torch.uint8),
                          rb.shadow.reshape(-1, 1).view(torch.uint8))]
    r4_ms = timed(render_and_gather, iters=5, warm=3)          # 3 warm-ups: the first 4K frames grow the allocator's pools
    render_only_ms = timed(lambda: renderer.render(net, ro, rd), iters=5, warm=1)
    out["render_4k_shadow"] = {"rays": w4 * h4, "rays_this_rank": int(ro.shape[0]), "ms": r4_ms, "fps": 1e3 / r4_ms,
                               "rays_per_s": w4 * h4 / (r4_ms / 1e3), "render_only_ms": render_only_ms,
                               "gathered_bytes": int((w4 * h4 - ro.shape[0]) * 18),
                               "note": "ONE 3840x2160 frame: primary trace + ground plane + shadow trace + normals "
                                       "(Renderer.render) over 4 interleaved column strips per rank, then the final gather of "
                                       "depth / normal / hit / shadow to rank 0 (NCCL gather) inside the timed region"}
    if rank == 0 and frame.get("full") is not None and frame["full"][2] is not None:
        out["render_4k_shadow"]["gathered_frame_hits"] = int(frame["full"][2].sum())
        out["render_4k_shadow"]["gathered_frame_shadow_pixels"] = int(frame["full"][3].sum())
    del ro, rd, renderer, frame

    # ---- (b) config 3: 500 000 points per step over the ranks
    V, F = normalize(*[t.to(device) for t in torus(0.6, 0.25, 128, 64)])
    per_rank = 500000 // world
    modes = ["rand", "near", "near", "trace", "trace"]
    tri = V[F].contiguous()

    def make_batch():
        pts = point_sample(V, F, modes, per_rank // 5)
        return pts, ops.mesh2sdf_gpu(pts, tri)[0].unsqueeze(1)
    sample_ms = timed(make_batch, iters=5, warm=2)
    tnet = copy.deepcopy(net)
    tnet.train()
    trainer = FusedTrainer(tnet, lr=1e-3)
    pts, gts = make_batch()
    step_ms = timed(lambda: trainer.step(pts, gts, global_batch=per_rank * world), iters=5, warm=2)
    # the whole config-3 step with the next batch produced on a second stream while this one trains
    side = torch.cuda.Stream(device)
    state = {"batch": make_batch()}

    def pipelined_step():
        cur = torch.cuda.current_stream(device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            nxt = make_batch()
        p, g_ = state["batch"]
        trainer.step(p, g_, global_batch=per_rank * world)
        cur.wait_stream(side)
        state["batch"] = nxt
    pipe_ms = timed(pipelined_step, iters=6, warm=2)
    comm = {}
    if world > 1:
        dist = torch.distributed
        comm["reduce_scatter_plus_all_gather_ms"] = timed(lambda: (
            dist.reduce_scatter_tensor(trainer.grad_shard, trainer.flat_grad, op=dist.ReduceOp.SUM),
            dist.all_gather_into_tensor(trainer.flat_grad, trainer.grad_shard)), iters=5, warm=2)
        comm["all_reduce_ms"] = timed(lambda: dist.all_reduce(trainer.flat_grad), iters=5, warm=2)
        comm["bytes"] = int(trainer.flat_grad.numel() * 4)
    out["train_step_500k"] = {"points_per_rank": per_rank, "step_ms": step_ms, "sample_and_label_ms": sample_ms,
                              "config3_step_ms_serial": step_ms + sample_ms, "config3_step_ms": pipe_ms,
                              "config3_points_per_s": per_rank * world / (pipe_ms / 1e3),
                              "optimizer": "sharded (reduce-scatter, Adam on 1/N, all-gather)" if trainer.sharded else "replicated",
                              "comm": comm, "mesh_triangles": int(tri.shape[0]),
                              "note": "step_ms: fused fwd+loss+bwd for 5 LODs + gradient exchange + Adam on a resident batch; "
                                      "config3_step_ms: the same with a FRESH batch per step (sampler kernel + mesh2sdf labels) "
                                      "produced on a second stream while the previous batch trains"}
    log(f"strong scaling @N={world}: 4K+shadow+gather {r4_ms:.2f} ms (render {render_only_ms:.2f}), train step {step_ms:.2f} ms, "
        f"sample+label {sample_ms:.2f} ms, config-3 pipelined {pipe_ms:.2f} ms, comm {comm}")
    return out


# ----------------------------------------------------------------------------------------------- other configs
def run_extras(net, args, device, rank, world, flush, log):
    """Side measurements (device-timed, L2 flushed, a few iterations each).  They are reported under "extras"; the headline
    value / roofline above are untouched by them."""
    from nglod_b200 import dist as ndist
    from nglod_b200.lib import spc as S
    from nglod_b200.lib.renderer import Renderer
    from nglod_b200.lib.tracer import SphereTracer
    from nglod_b200.lib.torchgp import torus, normalize
    from nglod_b200.lib.geoutils import look_at
    out = {}
    import copy

    def timed(fn, iters=5, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            flush_l2(flush)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return ndist.max_over_ranks(float(np.mean(ts)), device)

    # ---- config 2, whole user-facing frame: Renderer.shade_images (rays from look_at + trace + matcap shading on the
    #      device + every buffer copied to the host and transposed to (H,W,C)), wall clock
    rargs2 = copy.copy(args)
    rargs2.render_res = [W, H]
    r2 = Renderer(SphereTracer(rargs2), args=rargs2, device=device)
    img = None
    for _ in range(5):      # the first calls page-lock their host buffers (~100 ms per new 52 MB set); steady state reuses them
        img = r2.shade_images(net, f=CAM_FROM, t=CAM_TO, fov=FOV)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        img = r2.shade_images(net, f=CAM_FROM, t=CAM_TO, fov=FOV)
    shade_s = ndist.max_over_ranks((time.perf_counter() - t0) / 10, device)
    out["shade_images_720p"] = {"ms": shade_s * 1e3, "fps": 1.0 / shade_s, "hit_pixels": int(img.hit.sum()),
                                "note": "Renderer.shade_images end to end, steady state (after the host buffers are page-locked): look_at "
                                        "rays, sphere trace, matcap shading, all RenderBuffer fields to host, (H,W,C) layout"}

    # ---- config 4 as BASELINE.json states it: 1920x1080, octree level 7, a 6-LOD natively sparse model (base-lod 2, so LOD l
    #      lives on level l + 2), lods 1..5: traversal + first voxel + in-voxel tracing with re-location + normals
    V, F = normalize(*[t.to(device) for t in torus(0.6, 0.25, 128, 64)])
    torch.manual_seed(77 + rank)
    octree = S.mesh_to_octree(V, F, 7, num_samples=1 << 22)
    spc = S.SPC(octree)
    ro, rd = look_at(CAM_FROM, CAM_TO, 1920, 1080, mode="persp", fov=FOV, device=device)
    nug = spc.raytrace(ro, rd, 7)
    trav_ms = timed(lambda: spc.raytrace(ro, rd, 7), iters=5, warm=1)
    out["spc_raytrace_1080p_level7"] = {"rays": ro.shape[0], "nuggets": int(nug.shape[0]), "voxels": int(spc.pyramid[0, 7]),
                                        "ms": trav_ms, "rays_per_s": world * ro.shape[0] / (trav_ms / 1e3),
                                        "note": "count + scan + fill, includes the one 4-byte host read of the total"}
    del nug
    nspc = S.NeuralSPC(spc, num_lods=6, base_lod=2)
    # a short in-run fit of the sparse model (fused sparse step, all six heads, analytic torus labels inside occupied voxels)
    opt = torch.optim.Adam(nspc.parameters(), lr=1e-3)
    gs = torch.Generator(device=device).manual_seed(11 + rank)
    lp7 = spc.level_points(7)[:, :3].float()
    for it in range(120):
        pv = torch.randint(0, lp7.shape[0], (65536,), device=device, generator=gs)
        xs7 = ((lp7[pv] + torch.rand(pv.shape[0], 3, device=device, generator=gs)) / 128 * 2 - 1).contiguous()
        gt7 = (torch.sqrt((torch.sqrt(xs7[:, 0] ** 2 + xs7[:, 2] ** 2) - 0.6) ** 2 + xs7[:, 1] ** 2) - 0.25).unsqueeze(1)
        opt.zero_grad(set_to_none=False)
        nspc.loss_backward(xs7, gt7)
        opt.step()
    nspc.eval()
    sweep = {}
    with torch.no_grad():
        for lod in (1, 2, 3, 4, 5):
            st = torch.zeros(2, dtype=torch.int64, device=device)
            x_, t_, hit_, n_, p_ = nspc.trace(ro, rd, lod, stats=st)
            ms = timed(lambda: nspc.trace(ro, rd, lod), iters=3, warm=1)
            sweep[f"lod{lod}"] = {"ms": ms, "rays_per_s": world * ro.shape[0] / (ms / 1e3), "hits": int(hit_.sum()),
                                  "octree_level": lod + 2, "sdf_evals_per_ray": int(st[0]) / ro.shape[0]}
    out["spc_sphere_trace_1080p"] = {"rays": ro.shape[0], "octree_level": 7, "voxels": int(spc.pyramid[0, 7]), "num_lods": 6,
                                     "corner_rows": int(nspc.corner_feats.shape[0]), "lods": sweep,
                                     "note": "BASELINE configs[3]: NeuralSPC num_lods=6 over the level-7 octree, fitted in-run "
                                             "(120 fused sparse steps x 65 536 points x 6 heads); per LOD: traverse at level "
                                             "lod+2 + first voxel + in-voxel trace (50 steps, far 5) + normals, one host read "
                                             "per frame (nugget total)"}
    # ---- the sparse training step itself, 500 000 points in occupied level-7 voxels, head 5
    pv = torch.randint(0, lp7.shape[0], (500000 // world,), device=device, generator=gs)
    xs7 = ((lp7[pv] + torch.rand(pv.shape[0], 3, device=device, generator=gs)) / 128 * 2 - 1).contiguous()
    gt7 = (torch.sqrt((torch.sqrt(xs7[:, 0] ** 2 + xs7[:, 2] ** 2) - 0.6) ** 2 + xs7[:, 1] ** 2) - 0.25).unsqueeze(1)
    nspc.train()

    def sparse_step_fused():          # one kernel: forward + loss + backward of the head (NeuralSPC.loss_backward)
        opt.zero_grad(set_to_none=False)
        nspc.loss_backward(xs7, gt7, lods=[5], pidx=[pv])
        opt.step()
    spf_ms = timed(sparse_step_fused, iters=5, warm=2)
    out["neural_spc_train_step_500k"] = {"ms_per_step": spf_ms, "points_per_s": world * pv.shape[0] / (spf_ms / 1e3),
                                         "voxels_level7": int(lp7.shape[0]), "corner_rows": int(nspc.corner_feats.shape[0]),
                                         "note": "NeuralSPC (6 LODs, levels 2-7), head 5: fused sparse forward + loss + backward "
                                                 "through the parent chain (nglod_sparse_sdf_train_step) + torch Adam; no "
                                                 "gradient all-reduce"}
    del nspc, opt, ro, rd

    # ---- the streaming (HBM-bound) kernels around the hot path, each timed alone against the measured copy bandwidth:
    #      algorithmic bytes (every operand read once, every result written once) / launch time, L2 flushed before each launch
    from nglod_b200 import ops as _ops
    from nglod_b200.lib.geoutils import camera_basis, procedural_matcap
    hbm_peak, _, peak_src = load_peaks()
    stream = {}

    def stream_entry(name, nbytes, fn, what):
        ms = timed(fn, iters=7, warm=3)
        stream[name] = {"ms": ms, "bytes": int(nbytes), "achieved_GBs": nbytes / ms / 1e6, "peak_GBs": hbm_peak,
                        "frac": nbytes / ms / 1e6 / hbm_peak, "what": what}
    n_par = sum(p.numel() for p in net.parameters())
    flat = [torch.zeros(n_par, device=device) for _ in range(4)]
    flat[1].normal_()
    stream_entry("nglod_adam_step", 28 * n_par, lambda: _ops.adam_step(flat[0], flat[1], flat[2], flat[3], 1),
                 f"Adam over the {n_par} parameters of the model as one flat buffer: reads p, g, m, v, writes p, m, v")
    del flat
    o4, v4, r4, u4 = camera_basis(CAM_FROM, CAM_TO)
    wx4 = torch.linspace(-1, 1, 3840, device=device) * (3840 / 2160)
    wy4 = torch.linspace(1, -1, 2160, device=device)
    rays4 = (torch.empty(3840 * 2160, 3, device=device), torch.empty(3840 * 2160, 3, device=device))
    tan4 = np.float32(np.tan(np.radians(FOV / 2)))
    stream_entry("nglod_generate_rays", 24 * 3840 * 2160,
                 lambda: _ops.generate_rays(o4, v4, r4, u4, tan4, False, wx4, wy4, out=rays4), "3840x2160 rays: writes ray_o, ray_d")
    nrm4 = torch.nn.functional.normalize(torch.randn(3840 * 2160, 3, device=device), dim=-1)
    hit4 = torch.rand(3840 * 2160, device=device) < 0.5
    mat = procedural_matcap(device=device).tex.float().contiguous()
    stream_entry("nglod_shade_matcap", (12 + 12 + 1 + 12 + 6) * 3840 * 2160, lambda: _ops.shade_matcap(rays4[1], nrm4, hit4, mat),
                 "3840x2160 pixels, half of them hits: reads view, normal, hit, writes rgb and the normals of the misses")
    del rays4, nrm4, hit4
    top = net.features[LOD].fm.data
    view4 = net.net_view()
    half_out = _ops.pack_grid_fp16(view4.summed[LOD])
    stream_entry("nglod_pack_grid_fp16", top.numel() * 6, lambda: _ops.pack_grid_fp16(view4.summed[LOD], out=half_out),
                 "the 65^3 x 32 prefix-summed grid to fp16 x-pair lines: reads fp32, writes fp16")
    del half_out
    sum_out = torch.empty_like(view4.summed[LOD])
    coarse = net.features[LOD - 1].fm.data.numel()
    stream_entry("nglod_build_summed_grid", top.numel() * 8 + coarse * 4,
                 lambda: _ops.build_summed_grid(view4, LOD, out=sum_out),
                 "level 4 of the prefix-summed grid = prolongation of level 3 + grid 4: reads both, writes 65^3 x 32 fp32")
    del sum_out
    out["streaming_kernels"] = dict(stream, note="HBM-bound kernels of the path (optimiser, ray generation, shading, derived-grid "
                                                 "builds); peak = MEASURED_PEAKS.json hbm_gbs (" + peak_src + ")")
    log("extras: streaming kernels " + ", ".join(f"{k} {v['frac']:.2f}" for k, v in stream.items()))

    # ---- the headless real-time loop (ray generation -> trace -> matcap shading, frame stays on the device)
    from nglod_b200.app import realtime
    rt = {}
    for tag, kw in (("dense_1080p_lod4", dict(lod=LOD)), ("sparse_level6_1080p_lod4", dict(lod=LOD, spc_level=6))):
        r = realtime.run(net, 1920, 1080, frames=24, **kw)
        steady = r["ms"][3:]
        rt[tag] = {"ms_per_frame": ndist.max_over_ranks(float(np.mean(steady)), device), "fps": 1e3 / float(np.mean(steady)),
                   "hit_pixels_last_frame": int(r["hit"].sum())}
    out["realtime_loop"] = dict(rt, note="app/realtime.py: orbiting camera, nglod_generate_rays -> tracer -> "
                                         "nglod_shade_matcap into a device RGB buffer; CUDA-event time per frame, no L2 flush")
    log(f"extras: spc traversal 1080p/level 7 {trav_ms:.2f} ms, sparse trace lods 1-5 "
        f"{[round(v['ms'], 2) for v in sweep.values()]} ms, sparse step {spf_ms:.2f} ms")
    return out


REF_GPU_SCRIPT = r'''
import json, sys, time, numpy as np, torch
sys.path.insert(0, %(root)r)
from oracle import ref_python
ref = ref_python.import_reference("reference")
dev = "cuda"
args = ref.parse_options(return_parser=True).parse_args(["--net", "OctreeSDF", "--num-lods", "5", "--feature-dim", "32", "--lod", "4"])
net = ref.OctreeSDF(args)
net.load_state_dict(torch.load(%(weights)r, map_location="cpu"))
net = net.to(dev).eval(); net.lod = 4
tracer = ref.SphereTracer(args)
torch.manual_seed(1000)
o, d = ref.look_at(%(cam)r, [0.0, 0.0, 0.0], 1280, 720, fov=30.0, mode="persp", device="cpu")
o, d = o.to(dev), d.to(dev)
def timed(fn, it, warm):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it
res = {}
with torch.no_grad():
    rb = tracer(net, o, d)
    res["trace_720p_ms"] = timed(lambda: tracer(net, o, d), 5, 2)
    res["hits"] = int(rb.hit.sum())
g = torch.Generator(device=dev).manual_seed(1)
x = torch.rand(1 << 20, 3, device=dev, generator=g) * 2 - 1
gt = torch.rand(1 << 20, 1, device=dev, generator=g)
with torch.no_grad():
    res["sdf_forward_2^20_ms"] = timed(lambda: net.sdf(x, lod=4), 20, 5)
def fb():
    net.zero_grad(set_to_none=True)
    (((net.sdf(x, lod=4) - gt) ** 2).sum() / x.shape[0]).backward()
res["sdf_forward_backward_2^20_ms"] = timed(fb, 10, 3)
import mesh2sdf
sys.path.insert(0, %(root)r)
from nglod_b200.lib.torchgp import torus
V, F = ref.torchgp.normalize(*torus(0.6, 0.25, 128, 64))
tri = V[F].to(dev).contiguous()
pts = ref.torchgp.point_sample(V, F, ["rand", "near", "near", "trace", "trace"], 100000).to(dev).contiguous()
res["mesh2sdf_500k_x_16k_ms"] = timed(lambda: mesh2sdf.mesh2sdf_gpu(pts, tri), 3, 1)
print("RESULT " + json.dumps(res))
'''


def reference_on_this_gpu(net, log):
    """SURVEY 8d's second baseline: the UNMODIFIED reference (its Python from oracle/_ref/sdf-net, its own two CUDA
    extensions from oracle/_ref) on THIS B200, same weights, same rays, CUDA-event timed -- in a subprocess, because the
    reference binds its extension modules at import.  Reported next to ours; not a target."""
    import tempfile
    try:
        with tempfile.TemporaryDirectory() as td:
            wpath = os.path.join(td, "w.pt")
            torch.save({k: v.detach().cpu() for k, v in net.state_dict().items()}, wpath)
            script = REF_GPU_SCRIPT % {"root": ROOT, "weights": wpath, "cam": CAM_FROM}
            r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=600)
        if r.returncode != 0:
            return {"unavailable": r.stderr.strip().splitlines()[-1][:300] if r.stderr.strip() else "failed"}
        res = json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])
        res["rays_per_s"] = W * H / (res["trace_720p_ms"] / 1e3)
        res["sdf_forward_qps"] = SDF_N / (res["sdf_forward_2^20_ms"] / 1e3)
        res["sdf_forward_backward_qps"] = SDF_N / (res["sdf_forward_backward_2^20_ms"] / 1e3)
        res["note"] = ("the unmodified reference classes (oracle/_ref/sdf-net) + the reference's own compiled sol_nglod / mesh2sdf "
                       "kernels (oracle/_ref) on this GPU, same fitted weights and rays as ours; no L2 flush")
        log(f"reference on this GPU: trace {res['trace_720p_ms']:.1f} ms, sdf fwd {res['sdf_forward_2^20_ms']:.2f} ms, "
            f"fwd+bwd {res['sdf_forward_backward_2^20_ms']:.2f} ms, mesh2sdf {res['mesh2sdf_500k_x_16k_ms']:.1f} ms")
        return res
    except Exception as e:  # noqa: BLE001
        return {"unavailable": str(e)[:300]}


# ----------------------------------------------------------------------------------------------- CPU legs
def subsample_rays(ray_o, ray_d, stride):
    """Every `stride`-th column and row of the x-major frame (ray = ix*H + iy)."""
    o = ray_o.reshape(W, H, 3)[::stride, ::stride].reshape(-1, 3).contiguous()
    d = ray_d.reshape(W, H, 3)[::stride, ::stride].reshape(-1, 3).contiguous()
    return o, d


def cpu_reference_tracer(state_dict, log):
    """(trace(o, d) -> RenderBuffer-like, kind): the UNMODIFIED reference's OctreeSDF + SphereTracer on the host cores
    (oracle/ref_python: reference Python from /root/reference or its staged copy under oracle/_ref, sol_nglod.aabb from
    oracle.c, which the GPU tests pin bit for bit to the reference's CUDA aabb).  Falls back to the oracle port."""
    torch.set_num_threads(os.cpu_count() or 1)
    try:
        from oracle import ref_python
        ref = ref_python.import_reference("cpu")
        rargs = ref.parse_options(return_parser=True).parse_args(
            ["--net", "OctreeSDF", "--num-lods", str(NUM_LODS), "--feature-dim", "32", "--lod", str(LOD)])
        rnet = ref.OctreeSDF(rargs)
        rnet.load_state_dict(state_dict)
        rnet.eval()
        rnet.lod = LOD
        tracer = ref.SphereTracer(rargs)

        def trace(o, d):
            with torch.no_grad():
                return tracer(rnet, o, d)
        return trace, "reference"
    except Exception as e:  # noqa: BLE001
        log(f"reference Python unavailable ({e}); timing the oracle port instead")
        from oracle import nglod_oracle as O
        onet = O.OracleNet(state_dict)
        onet.lod = LOD
        return (lambda o, d: O.sphere_trace(onet, o, d)), "port"


def cpu_trace_rate(trace, ray_o, ray_d, budget_s, log):
    """Time the CPU tracer on the largest sub-frame (every s-th row and column of the same rays) that fits the budget."""
    o, d = subsample_rays(ray_o, ray_d, 32)                 # 40x23 probe
    t0 = time.perf_counter()
    trace(o, d)
    probe = max(time.perf_counter() - t0, 1e-3)
    per_ray = probe / o.shape[0]
    stride = 32
    for s in (16, 8, 4, 2, 1):
        if per_ray * (W // s) * (H // s) * 0.6 <= budget_s:   # batches get more efficient as they grow
            stride = s
    o, d = subsample_rays(ray_o, ray_d, stride)
    t0 = time.perf_counter()
    trace(o, d)
    dt = time.perf_counter() - t0
    log(f"cpu tracer: {o.shape[0]} rays ({W // stride}x{H // stride}) in {dt:.2f}s on {torch.get_num_threads()} threads")
    return o.shape[0] / dt, stride, f"{W // stride}x{H // stride} sub-frame (every {stride}th row/col) of the same rays, {o.shape[0]} rays, 1 pass"


def cpu_baseline_trace(net, ray_o, ray_d, budget_s, log):
    trace, kind = cpu_reference_tracer({k: v.detach().cpu() for k, v in net.state_dict().items()}, log)
    rate, _, sample = cpu_trace_rate(trace, ray_o, ray_d, budget_s, log)
    return {"value": rate, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": kind, "sample": sample}


def run_reference(ns):
    """The reference arm: the UNMODIFIED reference's own CPU implementation of the path (its OctreeSDF + SphereTracer classes,
    PyTorch on the host cores) on the SAME fitted net and the SAME rays as our arm; each step a bounded sub-frame."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return

    def log(msg):
        print("[bench-ref] " + msg, file=sys.stderr, flush=True)

    torch.set_num_threads(os.cpu_count() or 1)
    # the parameter container comes from the reference's own class when its Python is available (same seed, same init order)
    try:
        from oracle import ref_python
        ref = ref_python.import_reference("cpu")
        rargs = ref.parse_options(return_parser=True).parse_args(["--net", "OctreeSDF", "--num-lods", str(NUM_LODS), "--feature-dim", "32"])
        torch.manual_seed(0)
        init = {k: v.detach().clone().contiguous() for k, v in ref.OctreeSDF(rargs).state_dict().items()}
        torch.manual_seed(1000)
        ray_o, ray_d = ref.look_at(camera_from(0.0), CAM_TO, W, H, mode="persp", fov=FOV, device="cpu")
    except Exception as e:  # noqa: BLE001
        log(f"reference Python unavailable ({e}); parameter container and rays from the package's mirror classes")
        from nglod_b200.lib.options import parse_options
        from nglod_b200.lib.models import OctreeSDF
        torch.manual_seed(0)
        init = {k: v.detach().clone().contiguous() for k, v in
                OctreeSDF(parse_options(return_parser=True).parse_args(["--net", "OctreeSDF", "--num-lods", str(NUM_LODS)])).state_dict().items()}
        ray_o, ray_d = make_rays("cpu")
    sd, how = fit_shared(init, log)
    trace, kind = cpu_reference_tracer(sd, log)
    total_steps = ns.steps + ns.warmup
    budget = float(os.environ.get("NGLOD_REF_BUDGET_S", max(2.0, min(20.0, 150.0 / max(total_steps, 1)))))
    _, stride, sample = cpu_trace_rate(trace, ray_o, ray_d, budget, log)
    o, d = subsample_rays(ray_o, ray_d, stride)
    for _ in range(ns.warmup):
        trace(o, d)
    t0 = time.perf_counter()
    for _ in range(ns.steps):
        trace(o, d)
    el = time.perf_counter() - t0
    value = o.shape[0] * ns.steps / el
    line = {
        "impl": "reference", "metric": "sphere_traced_rays_per_sec_1280x720_lod4", "value": value, "unit": "rays/s",
        "n_gpus": ns.gpus, "steps": ns.steps, "warmup": ns.warmup, "ms_per_step": el / ns.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": shared_config(ns.gpus),
        "details": {"fit_source": how},
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": kind,
                         "sample": sample + f", x{ns.steps} steps"},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the side measurements (sparse path, real-time loop, variants)")
    ns = ap.parse_args()
    # stdout carries exactly ONE line, the JSON: libraries that write to fd 1 (NCCL prints its version banner there)
    # are sent to stderr for the duration of the run; print() below still reaches the real stdout
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    if ns.impl == "reference":
        run_reference(ns)
    else:
        run_ours(ns)
    sys.stdout.flush()
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
