#!/usr/bin/env python
"""bench.py -- headline benchmark of the NGLOD hot path on N B200s (one process per GPU).

  python bench.py --gpus 1 --steps 20 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...      # the reference's CPU path (oracle port) on the host cores

Workload (BASELINE.json configs[1]): SphereTracer.forward over a 1280x720 perspective frame at lod 4 of a 5-LOD
OctreeSDF (feature-dim 32, hidden 128) fitted IN-RUN to a procedural torus mesh (mesh2sdf-labelled samples).
A "step" is one full frame (921 600 rays) per GPU; with N GPUs every rank traces its own frame (a different
camera azimuth), no data-path collective -> "scaling": "weak"; value = N * 921600 / max-over-ranks time.
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
FIT_STEPS, FIT_BATCH = 300, 65536
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
# dram__bytes_read.sum + dram__bytes_write.sum of ONE sphere_trace_kernel / sdf_forward_tc_kernel launch (single-grid
# fp32 instances), from the `ncu --set full` captures summarised in profiles/sphere_trace_sum_r1_v4.txt and
# profiles/sdf_forward_sum_r1_v4.txt
NCU_TRAFFIC_TRACE_BYTES = 38498048 + 2459392
NCU_TRAFFIC_SDF_FWD_BYTES = 47828992 + 575744


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), "measured"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback"


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


# ----------------------------------------------------------------------------------------------- set-up
def make_rays(device, azimuth_deg=0.0, seed=0):
    """look_at rays for the 720p frame; the camera is rotated about y by `azimuth_deg` (rank-dependent)."""
    from nglod_b200.lib.geoutils import look_at
    a = math.radians(azimuth_deg)
    f = [CAM_FROM[0] * math.cos(a) + CAM_FROM[2] * math.sin(a), CAM_FROM[1],
         -CAM_FROM[0] * math.sin(a) + CAM_FROM[2] * math.cos(a)]
    torch.manual_seed(1000 + seed)
    return look_at(f, CAM_TO, W, H, mode="persp", fov=FOV, device=device)


def build_and_fit(device, log):
    """5-LOD OctreeSDF fitted in-run to a procedural torus mesh: samples from the MeshDataset recipe
    (rand/near/near/trace/trace), labels from the mesh2sdf kernel, fused training step + Adam kernel."""
    from nglod_b200.lib.options import parse_options
    from nglod_b200.lib.models import OctreeSDF
    from nglod_b200.lib.datasets import MeshDataset
    from nglod_b200.lib.trainer import FusedTrainer
    from nglod_b200.lib.torchgp import torus
    args = parse_options(return_parser=True).parse_args(
        ["--net", "OctreeSDF", "--num-lods", str(NUM_LODS), "--feature-dim", "32", "--lod", str(LOD),
         "--render-res", str(W), str(H)])
    torch.manual_seed(0)
    net = OctreeSDF(args).to(device)
    net.math_mode = MATH_MODE
    net.grid_storage, net.sum_lods = GRID_STORAGE, SUM_LODS
    t0 = time.time()
    ds = MeshDataset(args, mesh=torus(0.6, 0.25, 128, 64), device=device)       # 500 000 labelled points
    torch.cuda.synchronize()
    t_ds = time.time() - t0
    trainer = FusedTrainer(net, lr=1e-3)
    n = len(ds)
    g = torch.Generator(device=device).manual_seed(7)
    t0 = time.time()
    last = None
    for it in range(FIT_STEPS):
        idx = torch.randint(0, n, (FIT_BATCH,), device=device, generator=g)
        last = trainer.step(ds.pts[idx], ds.d[idx])
    torch.cuda.synchronize()
    log(f"fit: dataset {n} pts in {t_ds:.2f}s, {FIT_STEPS} steps x {FIT_BATCH} pts in {time.time() - t0:.2f}s, "
        f"final loss {float(last):.3e}")
    net.lod = LOD
    net.eval()
    return net, args


def flush_l2(buf):
    buf.zero_()


# ----------------------------------------------------------------------------------------------- ours
def run_ours(ns):
    from nglod_b200 import dist as ndist
    from nglod_b200 import ops
    from nglod_b200.lib.tracer import SphereTracer
    rank, world, local = ndist.init_from_env()
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)

    def log(msg):
        if rank == 0:
            print("[bench] " + msg, file=sys.stderr, flush=True)

    net, args = build_and_fit(device, log)
    if world > 1:                       # identical weights on every rank
        for p in net.parameters():
            torch.distributed.broadcast(p.data, src=0)
    tracer = SphereTracer(args)
    ray_o, ray_d = make_rays(device, azimuth_deg=360.0 * rank / max(world, 1), seed=rank)
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

    # ---- e2e: same frame through the public API with HOST buffers (pinned), copies inside the timed region
    ho, hd = ray_o.cpu().pin_memory(), ray_d.cpu().pin_memory()
    out_host = {k: torch.empty(s, dtype=dt).pin_memory() for k, s, dt in
                (("x", (n_rays, 3), torch.float32), ("depth", (n_rays, 1), torch.float32),
                 ("hit", (n_rays,), torch.bool), ("normal", (n_rays, 3), torch.float32))}

    def e2e_step():
        # the user-facing host call: pinned rays in, pinned RenderBuffer fields out, copies pipelined with the trace
        tracer.trace_host(net, ho, hd, out=out_host)

    for _ in range(3):
        e2e_step()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(ns.steps):
        e2e_step()
    e2e_s = ndist.max_over_ranks(time.perf_counter() - t0, device)
    clock_info = clocks.stop()

    # ---- SDF query throughput (BASELINE.json configs[0]): 2^20 random points, lod 4, forward and forward+backward
    g = torch.Generator(device=device).manual_seed(1)
    xq = torch.rand(SDF_N, 3, device=device, generator=g) * 2 - 1
    gq = torch.rand(SDF_N, device=device, generator=g)

    def time_kernel(fn, iters):
        for _ in range(3):
            flush_l2(flush)
            fn()
        torch.cuda.synchronize()
        ev = []
        for _ in range(iters):
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
    extras = {} if ns.no_extras else run_extras(net, args, device, rank, world, flush, log)
    if variants:
        extras["inference_variants"] = variants
    if rank != 0:
        return
    peak, peak_src = load_peaks()
    value = world * n_rays * ns.steps / (total_ms / 1e3)
    alg_bytes = n_eval * (GATHER_BYTES_PER_QUERY + IO_BYTES_PER_QUERY) + n_rays * RAY_IO_BYTES
    achieved = alg_bytes / (kernel_ms / 1e3) / 1e9
    compulsory = (n_rays * RAY_IO_BYTES + sum(f.fm.numel() * 4 for f in net.features)) / (kernel_ms / 1e3) / 1e9
    q_bytes = SDF_N * (GATHER_BYTES_PER_QUERY + IO_BYTES_PER_QUERY)
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
        "config": {"workload": "SphereTracer.forward 1280x720 persp fov30 lod4, OctreeSDF num-lods=5 feature-dim=32 "
                               "hidden=128 fitted in-run to a procedural torus mesh (BASELINE.json configs[1])",
                   "rays_per_gpu_step": n_rays, "fps_per_gpu": 1e3 / (total_ms / ns.steps),
                   "sdf_evals_per_ray": n_eval / n_rays, "hit_fraction": n_hit / n_rays,
                   "num_steps": 256, "math_mode": MATH_MODE, "grid_storage": GRID_STORAGE,
                   "lod_sum": "prefix-summed grid (one 8-corner gather per evaluation)" if SUM_LODS else "per-LOD gather",
                   "l2": "flushed between timed iterations (256 MB memset)",
                   "parallelism": f"one frame per GPU x{world}, no collective"},
        "e2e": {"value": world * n_rays * ns.steps / e2e_s, "unit": "rays/s",
                "h2d_bytes_per_step": n_rays * 24, "d2h_bytes_per_step": n_rays * (12 + 4 + 1 + 12),
                "ms_per_step": e2e_s / ns.steps * 1e3},
        "gpu_launches": ns.steps,
        "clocks": clock_info,
        "roofline": {"bound": "hbm", "kernel": "sphere_trace_kernel", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": NCU_TRAFFIC_TRACE_BYTES, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kernel_ms,
                     "compulsory_GBs": compulsory,
                     "per_lod_formulation_GBs": (n_eval * (GATHER_BYTES_PER_QUERY_PER_LOD + IO_BYTES_PER_QUERY)
                                                 + n_rays * RAY_IO_BYTES) / (kernel_ms / 1e3) / 1e9,
                     "note": f"algorithmic bytes = sdf_evals x ({GATHER_BYTES_PER_QUERY} B gather + 16 B io) + rays x 53 B; "
                             "per_lod_formulation_GBs is the same count with SURVEY 8d's 5120 B (the reference's 5 "
                             "separate gathers, which the summed grid replaces); the gather is "
                             "served by L2/L1 (summed grid 35 MB < L2), so frac may exceed 1 (SURVEY.md 8d); DRAM traffic "
                             "(ncu) is the compulsory ~40 MB: the grid once + ray I/O. The kernel is latency/sync "
                             "bound (issue 37 %, L1 34 %, L2 8 %, tensor 24 %), see DESIGN.md section 4"},
        "sdf_queries": {"n": SDF_N, "lod": LOD,
                        "forward_qps": world * SDF_N / (fwd_ms / 1e3), "forward_ms": fwd_ms,
                        "forward_backward_qps": world * SDF_N / ((fwd_ms + bwd_ms) / 1e3), "backward_ms": bwd_ms,
                        "roofline": {"bound": "hbm", "kernel": "sdf_forward_tc_kernel" if MATH_MODE == "tc" else "sdf_forward_kernel",
                                     "achieved": q_bytes / (fwd_ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                                     "frac": q_bytes / (fwd_ms / 1e3) / 1e9 / peak,
                                     "traffic": NCU_TRAFFIC_SDF_FWD_BYTES,
                                     "l2_to_sm_GBs": SDF_N * GATHER_BYTES_PER_QUERY / (fwd_ms / 1e3) / 1e9,
                                     "note": "random queries miss L1 (hit 3 %), so the 1024 B/query gather is L2->SM "
                                             "traffic; the L2 fabric tops out near 6300 B/clk (~11 TB/s at 1.8 GHz)"}},
    }
    line["extras"] = extras
    if world == 1 and not ns.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_trace(net, ray_o.cpu(), ray_d.cpu(), budget_s=20.0, log=log)
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- other configs
def run_extras(net, args, device, rank, world, flush, log):
    """Side measurements for BASELINE.json configs 3-5 (device-timed, L2 flushed, a few iterations each).  They are
    reported under "extras"; the headline value / roofline above are untouched by them."""
    from nglod_b200 import dist as ndist
    from nglod_b200 import ops
    from nglod_b200.lib import spc as S
    from nglod_b200.lib.datasets import MeshDataset
    from nglod_b200.lib.renderer import Renderer
    from nglod_b200.lib.tracer import SphereTracer
    from nglod_b200.lib.trainer import FusedTrainer
    from nglod_b200.lib.torchgp import torus, point_sample, normalize
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

    # ---- config 3: training step on a 500 000-point batch sharded over the ranks (fused 5-head fwd+bwd, one
    #      all-reduce of the flat gradient, Adam kernel) + the cost of producing the batch (sampling + mesh2sdf labels)
    V, F = normalize(*[t.to(device) for t in torus(0.6, 0.25, 128, 64)])
    per_rank = 500000 // world
    modes = ["rand", "near", "near", "trace", "trace"]
    tri = V[F].contiguous()

    def make_batch():
        pts = point_sample(V, F, modes, per_rank // 5)
        return pts, ops.mesh2sdf_gpu(pts, tri)[0].unsqueeze(1)

    sample_ms = timed(make_batch, iters=5, warm=2)
    pts, gts = make_batch()
    tnet = copy.deepcopy(net)
    tnet.train()
    trainer = FusedTrainer(tnet, lr=1e-3)
    step_ms = timed(lambda: trainer.step(pts, gts, global_batch=per_rank * world), iters=5, warm=2)
    out["train_step_500k"] = {"points_per_rank": pts.shape[0], "ms_per_step": step_ms,
                              "points_per_s": world * pts.shape[0] / (step_ms / 1e3),
                              "sample_and_label_ms": sample_ms, "mesh_triangles": int(tri.shape[0]),
                              "mesh2sdf_pairs_per_s": world * pts.shape[0] * tri.shape[0] / (sample_ms / 1e3),
                              "config3_step_ms": step_ms + sample_ms,
                              "config3_points_per_s": world * pts.shape[0] / ((step_ms + sample_ms) / 1e3),
                              "note": "fused fwd+loss+bwd for 5 LODs + flat-gradient all-reduce + Adam; a fresh batch per step "
                                      "(SURVEY 8d config 3) adds sample_and_label_ms: the sampler kernel + mesh2sdf labels; "
                                      "mesh2sdf_pairs_per_s counts all N x T pairs although the large-batch path visits few of them"}
    del trainer, tnet

    # ---- config 4 (traversal half): sparse-octree ray traversal at 1920x1080, octree level 7 of the same mesh
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

    # ---- config 4 (tracing half): sparse OctreeSDF of the fitted net over the level-6 octree, in-voxel sphere tracing
    #      with voxel re-location, lods 1-4 (lod l <-> octree level l+2), 1920x1080
    octree6 = S.mesh_to_octree(V, F, 6, num_samples=1 << 22)
    sp = S.SparseOctreeSDF(net, S.SPC(octree6))
    sweep = {}
    for lod in (1, 2, 3, 4):
        st = torch.zeros(2, dtype=torch.int64, device=device)
        x_, t_, hit_, n_, p_ = sp.trace(ro, rd, lod, stats=st)
        ms = timed(lambda: sp.trace(ro, rd, lod), iters=3, warm=1)
        sweep[f"lod{lod}"] = {"ms": ms, "rays_per_s": world * ro.shape[0] / (ms / 1e3), "hits": int(hit_.sum()),
                              "sdf_evals_per_ray": int(st[0]) / ro.shape[0]}
    out["spc_sphere_trace_1080p"] = {"rays": ro.shape[0], "octree_level": 6, "voxels": int(sp.spc.pyramid[0, 6]),
                                     "corner_rows": int(sp.corner_feats.shape[0]), "lods": sweep,
                                     "note": "traverse + first voxel + in-voxel trace (50 steps, far 5) + normals, one "
                                             "host read per frame (nugget total)"}
    del ro, rd, sp

    # ---- config 5: 3840x2160, shadows + normals, image cut into column strips over the ranks (x-major rays)
    w4, h4 = 3840, 2160
    torch.manual_seed(5)
    ro, rd = look_at(CAM_FROM, CAM_TO, w4, h4, mode="persp", fov=FOV, device=device)
    s0, s1 = ndist.shard_range(w4 * h4, rank, world, align=h4)
    ro, rd = ro[s0:s1].contiguous(), rd[s0:s1].contiguous()
    rargs = copy.copy(args)
    rargs.shadow, rargs.ground_height, rargs.render_res = True, -0.4, [(s1 - s0) // h4, h4]
    renderer = Renderer(SphereTracer(rargs), args=rargs, device=device)
    r4_ms = timed(lambda: renderer.render(net, ro, rd), iters=3, warm=1)
    out["render_4k_shadow"] = {"rays": w4 * h4, "rays_this_rank": s1 - s0, "ms": r4_ms, "fps": 1e3 / r4_ms,
                               "note": "primary trace + ground plane + shadow trace + normals via Renderer.render; "
                                       "ranks take contiguous column strips, no collective"}
    # ---- a natively sparse model (features on the corners of the level-7 octree only: 128^3 has no dense grid
    #      here) trained with autograd through the sparse kernels + torch Adam, 500 000 points inside occupied voxels
    nspc = S.NeuralSPC(spc, num_lods=6, base_lod=2)
    opt = torch.optim.Adam(nspc.parameters(), lr=1e-3)
    lp7 = spc.level_points(7)[:, :3].float()
    gs = torch.Generator(device=device).manual_seed(11 + rank)
    pv = torch.randint(0, lp7.shape[0], (500000 // world,), device=device, generator=gs)
    xs7 = ((lp7[pv] + torch.rand(pv.shape[0], 3, device=device, generator=gs)) / 128 * 2 - 1).contiguous()
    gt7 = (torch.sqrt((torch.sqrt(xs7[:, 0] ** 2 + xs7[:, 2] ** 2) - 0.6) ** 2 + xs7[:, 1] ** 2) - 0.25).unsqueeze(1)

    def sparse_step():
        opt.zero_grad(set_to_none=True)
        loss = ((nspc.sdf(xs7, 5, pv) - gt7) ** 2).mean()
        loss.backward()
        opt.step()
    sp_ms = timed(sparse_step, iters=5, warm=2)

    def sparse_step_fused():          # one kernel: forward + loss + backward of the head (NeuralSPC.loss_backward)
        opt.zero_grad(set_to_none=False)
        nspc.loss_backward(xs7, gt7, lods=[5], pidx=[pv])
        opt.step()
    spf_ms = timed(sparse_step_fused, iters=5, warm=2)
    out["neural_spc_train_step_500k"] = {"ms_per_step": spf_ms, "points_per_s": world * pv.shape[0] / (spf_ms / 1e3),
                                         "autograd_ms_per_step": sp_ms,
                                         "voxels_level7": int(lp7.shape[0]), "corner_rows": int(nspc.corner_feats.shape[0]),
                                         "note": "NeuralSPC (6 LODs, levels 2-7), head 5: fused sparse forward + loss + gen-2 "
                                                 "backward through the parent chain (nglod_sparse_sdf_train_step) + torch Adam; "
                                                 "autograd_ms_per_step = the same step as separate forward / loss / backward; "
                                                 "no gradient all-reduce"}
    del nspc, opt

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
    log(f"extras: train {step_ms:.2f} ms/500k-pt step, sample+label {sample_ms:.1f} ms, spc 1080p {trav_ms:.2f} ms, "
        f"4K+shadow {r4_ms:.1f} ms")
    return out


# ----------------------------------------------------------------------------------------------- CPU legs
def subsample_rays(ray_o, ray_d, stride):
    """Every `stride`-th column and row of the x-major frame (ray = ix*H + iy)."""
    o = ray_o.reshape(W, H, 3)[::stride, ::stride].reshape(-1, 3).contiguous()
    d = ray_d.reshape(W, H, 3)[::stride, ::stride].reshape(-1, 3).contiguous()
    return o, d


def cpu_trace_rate(onet, ray_o, ray_d, budget_s, log):
    """Time the oracle's batch-loop tracer on the host; picks the largest sub-frame that fits the budget."""
    from oracle import nglod_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    o, d = subsample_rays(ray_o, ray_d, 32)                 # 40x23 probe
    t0 = time.perf_counter()
    O.sphere_trace(onet, o, d)
    probe = max(time.perf_counter() - t0, 1e-3)
    per_ray = probe / o.shape[0]
    stride = 32
    for s in (16, 8, 4, 2, 1):
        if per_ray * (W // s) * (H // s) * 0.6 <= budget_s:   # batches get more efficient as they grow
            stride = s
    o, d = subsample_rays(ray_o, ray_d, stride)
    cnt = {}
    t0 = time.perf_counter()
    O.sphere_trace(onet, o, d, count=cnt)
    dt = time.perf_counter() - t0
    log(f"cpu tracer: {o.shape[0]} rays ({W // stride}x{H // stride}) in {dt:.2f}s on {torch.get_num_threads()} threads")
    return o.shape[0] / dt, f"{W // stride}x{H // stride} sub-frame (every {stride}th row/col) of the same rays, {o.shape[0]} rays, 1 pass", dt


def cpu_baseline_trace(net, ray_o, ray_d, budget_s, log):
    from oracle import nglod_oracle as O
    onet = O.OracleNet({k: v.detach().cpu() for k, v in net.state_dict().items()})
    onet.lod = LOD
    rate, sample, _ = cpu_trace_rate(onet, ray_o, ray_d, budget_s, log)
    return {"value": rate, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample}


def run_reference(ns):
    """The reference's CPU path for the same metric/config: the oracle port (same ATen calls as the reference's
    PyTorch path; the reference itself is Python under /root/reference, which does not exist on the GPU box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import nglod_oracle as O

    def log(msg):
        print("[bench-ref] " + msg, file=sys.stderr, flush=True)

    torch.set_num_threads(os.cpu_count() or 1)
    fit_dev = "cuda" if torch.cuda.is_available() else "cpu"
    # untimed set-up: fit the same architecture to the same torus with plain torch autograd (on the GPU if present)
    from nglod_b200.lib.options import parse_options
    from nglod_b200.lib.models import OctreeSDF
    args = parse_options(return_parser=True).parse_args(["--net", "OctreeSDF", "--num-lods", str(NUM_LODS)])
    torch.manual_seed(0)
    sd = OctreeSDF(args).state_dict()                     # parameter container only; no kernel is touched
    onet = O.OracleNet(sd, device=fit_dev, requires_grad=True)
    opt = torch.optim.Adam(onet.parameters(), lr=1e-3)
    g = torch.Generator(device=fit_dev).manual_seed(7)
    steps, batch = (FIT_STEPS, FIT_BATCH) if fit_dev == "cuda" else (60, 8192)
    steps = int(os.environ.get("NGLOD_REF_FIT_STEPS", steps))       # tests shrink the (untimed) set-up
    for _ in range(steps):
        p = torch.rand(batch, 3, device=fit_dev, generator=g) * 2 - 1
        surf = p[: batch // 2]
        q = torch.sqrt(surf[:, 0] ** 2 + surf[:, 2] ** 2)
        # half of the batch is pulled onto / near the torus surface, like the near/trace sample modes
        ring = torch.stack([surf[:, 0] / q * 0.6, torch.zeros_like(q), surf[:, 2] / q * 0.6], dim=1)
        dirv = torch.nn.functional.normalize(surf - ring, dim=1)
        p = torch.cat([ring + dirv * (0.25 + 0.01 * torch.randn(batch // 2, 1, device=fit_dev, generator=g)), p[batch // 2:]])
        qq = torch.sqrt(p[:, 0] ** 2 + p[:, 2] ** 2) - 0.6
        gt = (torch.sqrt(qq * qq + p[:, 1] ** 2) - 0.25).unsqueeze(1)
        O.l2_loss_and_grads(onet, p, gt, list(range(NUM_LODS)))
        opt.step()
    cpu_net = O.OracleNet({f"features.{i}.fm": t.detach().cpu() for i, t in enumerate(onet.fm)} |
                          {f"louts.{i}.{k}": t.detach().cpu() for i, dct in enumerate(onet.dec)
                           for k, t in zip(("0.weight", "0.bias", "2.weight", "2.bias"), dct)})
    cpu_net.lod = LOD
    # rays: the oracle's own look_at (same camera, seeded jitter)
    torch.manual_seed(1000)
    ray_o, ray_d = O.look_at(CAM_FROM, CAM_TO, W, H, mode="persp", fov=FOV)
    total_steps = ns.steps + ns.warmup
    budget = float(os.environ.get("NGLOD_REF_BUDGET_S", max(2.0, min(20.0, 150.0 / max(total_steps, 1)))))
    rate, sample, dt = cpu_trace_rate(cpu_net, ray_o, ray_d, budget, log)
    o, d = None, None
    # timed: W warm-up + K steps of the bounded sample
    stride = W // int(sample.split("x")[0])
    o, d = subsample_rays(ray_o, ray_d, stride)
    for _ in range(ns.warmup):
        O.sphere_trace(cpu_net, o, d)
    t0 = time.perf_counter()
    for _ in range(ns.steps):
        O.sphere_trace(cpu_net, o, d)
    el = time.perf_counter() - t0
    value = o.shape[0] * ns.steps / el
    line = {
        "impl": "reference", "metric": "sphere_traced_rays_per_sec_1280x720_lod4", "value": value, "unit": "rays/s",
        "n_gpus": ns.gpus, "steps": ns.steps, "warmup": ns.warmup, "ms_per_step": el / ns.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "SphereTracer.forward 1280x720 persp fov30 lod4, OctreeSDF num-lods=5 feature-dim=32 "
                               "hidden=128 fitted in-run to a torus (BASELINE.json configs[1]); CPU: " + sample},
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
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
    ap.add_argument("--no-extras", action="store_true", help="skip the config 3/4/5 side measurements")
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
