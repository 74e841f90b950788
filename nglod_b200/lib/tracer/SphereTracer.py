"""SphereTracer -- the reference's tracer class (sdf-net/lib/tracer/SphereTracer.py:39-245).

`forward(net, ray_o, ray_d)` returns `RenderBuffer(x, depth, hit, normal)` with the
reference's exact semantics (see csrc/tracer.cu for the per-ray state machine), but
an OctreeSDF is traced by ONE persistent sm_100a kernel that evaluates the SDF
inline -- no per-step launches, no boolean-mask gathers, no `cond.any()` sync.
Any other callable `net` goes through the same algorithm expressed with torch
ops on the device (still using the aabb kernel), so analytic SDFs keep working.
"""
import numpy as np
import torch
import torch.nn.functional as F

from .BaseTracer import BaseTracer
from .RenderBuffer import RenderBuffer
from ..diffutils import gradient
from ... import ops


def _is_octree(net):
    return hasattr(net, "net_view") and hasattr(net, "num_lods")


def _trace_lod(net):
    lod = getattr(net, "lod", None)
    return net.num_lods - 1 if (lod is None or not 0 <= lod < net.num_lods) else lod


class SphereTracer(BaseTracer):

    def forward(self, net, ray_o, ray_d):
        if _is_octree(net) and self.grad_method == "finitediff" and getattr(net, "interpolate", None) is None:
            x, depth, hit, normal = ops.sphere_trace(
                net.net_view(), _trace_lod(net), ray_o, ray_d, num_steps=self.num_steps, step_size=self.step_size,
                min_dis=self.min_dis, far=self.camera_clamp[1])
            return RenderBuffer(x=x, depth=depth, hit=hit, normal=normal)
        return self._forward_generic(net, ray_o, ray_d, track_min=False)

    def trace_host(self, net, ray_o, ray_d, out=None, chunks=3, streams=2, fractions=None):
        """`forward` for rays that live in (pinned) HOST memory, results into (pinned) host memory: the frame is cut
        into `chunks` contiguous ray ranges and the host->device copy of range i+1, the trace of range i and the
        device->host copy of range i-1 run concurrently (PCIe is full duplex; the ranges are independent because
        rays are).  Consecutive ranges are traced on two alternating streams so that the next range's CTAs fill the
        SMs the previous launch leaves idle in its tail (its slowest rays): back-to-back launches on ONE stream cost
        +40 % (measured).  3 ranges on 2 streams measured best for a 720p frame (profiles/exp_e2e_chunks.py); `fractions`
        gives explicit cumulative split points instead of equal ranges.  Returns a RenderBuffer of CPU tensors (`out`, if given, is reused: a dict
        with pinned x [N,3], depth [N,1], hit [N] bool, normal [N,3]).  Synchronises before returning."""
        if ray_o.is_cuda or ray_d.is_cuda:
            raise RuntimeError("trace_host: ray_o / ray_d are expected in host memory (use forward() for device tensors)")
        n = ray_o.shape[0]
        ray_o = ray_o.contiguous().float()
        ray_d = ray_d.contiguous().float()

        def stage(ws, bounds, s_in):
            evs = []
            with torch.cuda.stream(s_in):
                for a, b in zip(bounds[:-1], bounds[1:]):
                    ws["o"][a:b].copy_(ray_o[a:b], non_blocking=True)
                    ws["d"][a:b].copy_(ray_d[a:b], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(s_in)
                    evs.append(ev)
            return evs
        return self._host_pipeline(net, n, stage, out, ("x", "depth", "hit", "normal"), chunks, streams, fractions)

    def trace_lookat_host(self, net, f, t, width, height, fov=30.0, mode="persp", window=None, out=None,
                          fields=("depth", "hit", "normal"), chunks=3, streams=2, packed=False):
        """The camera-driven host call: what `Renderer.render_lookat` + `.cpu()` amounts to for a user of the reference
        (renderer.py:89-107: rays come from `look_at(f, t, W, H, fov)`, never from host ray buffers).  Inputs are the
        camera pose (host floats) and the jittered window coordinates `window = (wx [W], wy [H])` in pinned host memory
        (None: drawn on the device as `look_at` does); the W*H rays are generated on the device (nglod_generate_rays,
        x-major like the reference), traced in `chunks` ranges on alternating streams, and the requested RenderBuffer
        `fields` (any of x, depth, hit, normal) are copied into pinned host memory as each range finishes.
        Per frame over PCIe: (W + H) * 4 bytes up, 17 B/ray down for the default fields (no ray upload, no `x`).
        packed=True: ONE tracer launch that writes every ray's record straight into a pinned host buffer as the ray
        retires (nglod_sphere_trace_packed) -- no device->host copy of the frame, no chunking.  Without "x" in `fields` the
        records are 16 bytes {depth, normal} + one hit byte per ray, both written to pinned host memory by the kernel
        (17 B/ray over PCIe, nothing to copy afterwards); with "x" they are 32 bytes {depth, normal, hit, x}.  Returns a RenderBuffer whose fields are VIEWS of the pinned
        buffers (`out`: a dict {"packed": [W*H, 4 or 8] fp32, "hit": [W*H] bool} of pinned tensors to re-use).  The call
        returns after the frame has arrived.  Measured per 720p frame (profiles/exp_e2e_packed.py): see DESIGN.md section 5;
        also tried there: copies by idle warps of the kernel itself, and DMA chunks released by the running kernel through
        stream wait values (a chunk of the image completes only when its longest ray does)."""
        import numpy as np
        from ..geoutils import _window, camera_basis
        dev = next(net.parameters()).device
        n = width * height
        origin, view, right, up = camera_basis(f, t)
        tan = np.float32(np.tan(np.radians(fov / 2)))

        if packed:
            if not (_is_octree(net) and self.grad_method == "finitediff" and getattr(net, "interpolate", None) is None):
                raise RuntimeError("trace_lookat_host: only the fused OctreeSDF tracer has a host path")
            ws = getattr(self, "_packed_ws", None)
            if ws is None or ws["n"] != n or ws["dev"] != dev or ws["wh"] != (width, height):
                ws = {"n": n, "dev": dev, "wh": (width, height), "rays": torch.empty(6 * n + width + height, device=dev),
                      "queue": torch.empty(1, dtype=torch.int32, device=dev)}
                self._packed_ws = ws
            with_x = "x" in fields
            out = {} if out is None else out
            if "packed" not in out:
                out["packed"] = torch.empty(n, 8 if with_x else 4, pin_memory=True)      # torch's caching host allocator
            if not with_x and "hit" not in out:
                out["hit"] = torch.empty(n, dtype=torch.bool, pin_memory=True)
            with torch.cuda.device(dev):
                wx, wy = _window(width, height, dev) if window is None else window
                ops.sphere_trace_camera(net.net_view(), _trace_lod(net), origin, view, right,
                                        up, tan, mode == "ortho", wx, wy, ws["rays"], out["packed"],
                                        hit=None if with_x else out["hit"],
                                        num_steps=self.num_steps, step_size=self.step_size, min_dis=self.min_dis,
                                        far=self.camera_clamp[1], queue=ws["queue"])
                torch.cuda.current_stream(dev).synchronize()
            return RenderBuffer(**ops.unpack_trace(out["packed"], None if with_x else out["hit"]))

        def stage(ws, bounds, s_in):
            with torch.cuda.stream(s_in):
                if window is None:
                    with torch.cuda.device(dev):
                        wx, wy = _window(width, height, dev)
                else:
                    if "wx" not in ws or ws["wx"].shape[0] != width or ws["wy"].shape[0] != height:
                        ws["wx"], ws["wy"] = torch.empty(width, device=dev), torch.empty(height, device=dev)
                    ws["wx"].copy_(window[0], non_blocking=True)
                    ws["wy"].copy_(window[1], non_blocking=True)
                    wx, wy = ws["wx"], ws["wy"]
                ops.generate_rays(origin, view, right, up, tan, mode == "ortho", wx, wy,
                                  out=(ws["o"], ws["d"]))
                ev = torch.cuda.Event()
                ev.record(s_in)
            return [ev] * (len(bounds) - 1)
        return self._host_pipeline(net, n, stage, out, tuple(fields), chunks, streams, None)

    def _host_pipeline(self, net, n, stage, out, fields, chunks, streams, fractions):
        """Shared body of trace_host / trace_lookat_host: `stage(ws, bounds, s_in)` puts the rays of every range into the
        device workspace (ws["o"], ws["d"]) on stream s_in and returns one event per range."""
        if not (_is_octree(net) and self.grad_method == "finitediff" and getattr(net, "interpolate", None) is None):
            raise RuntimeError("trace_host: only the fused OctreeSDF tracer has a pipelined host path")
        dev = next(net.parameters()).device
        shapes = {"x": ((n, 3), torch.float32), "depth": ((n, 1), torch.float32), "hit": ((n,), torch.bool),
                  "normal": ((n, 3), torch.float32)}
        for k in fields:
            if k not in shapes:
                raise ValueError(f"unknown RenderBuffer field {k!r}")
        if out is None:
            out = {k: torch.empty(shapes[k][0], dtype=shapes[k][1], pin_memory=True) for k in fields}
        ws = getattr(self, "_host_ws", None)
        if ws is None or ws["n"] != n or ws["dev"] != dev:
            ws = {"n": n, "dev": dev,
                  "o": torch.empty(n, 3, device=dev), "d": torch.empty(n, 3, device=dev),
                  "x": torch.empty(n, 3, device=dev), "depth": torch.empty(n, 1, device=dev),
                  "hit": torch.empty(n, dtype=torch.bool, device=dev), "normal": torch.empty(n, 3, device=dev),
                  "queue": torch.empty(max(chunks, 1), dtype=torch.int32, device=dev),
                  "s_in": torch.cuda.Stream(dev), "s_out": torch.cuda.Stream(dev),
                  "s_c": [torch.cuda.Stream(dev) for _ in range(4)]}
            self._host_ws = ws
        view, lod = net.net_view(), _trace_lod(net)
        cur = torch.cuda.current_stream(dev)
        s_in, s_out = ws["s_in"], ws["s_out"]
        s_in.wait_stream(cur)           # order after whatever the caller queued (e.g. a weight update)
        for sc in ws["s_c"]:
            sc.wait_stream(cur)
        chunks = max(1, min(chunks, n)) if n > 0 else 1
        if fractions is not None:       # explicit cumulative split points in (0, 1), e.g. (0.4, 0.7, 0.9)
            bounds = [0] + [int(n * f) for f in fractions] + [n]
            chunks = len(bounds) - 1
        else:
            bounds = [(n * i) // chunks for i in range(chunks + 1)]
        if ws["queue"].numel() < chunks:
            ws["queue"] = torch.empty(chunks, dtype=torch.int32, device=dev)
        ev_in = stage(ws, bounds, s_in)
        for i in range(chunks):
            a, b = bounds[i], bounds[i + 1]
            if b == a:
                continue
            sc = ws["s_c"][i % max(1, min(streams, 4))]
            sc.wait_event(ev_in[i])
            with torch.cuda.stream(sc):
                ops.sphere_trace(view, lod, ws["o"][a:b], ws["d"][a:b], num_steps=self.num_steps,
                                 step_size=self.step_size, min_dis=self.min_dis, far=self.camera_clamp[1],
                                 out=(ws["x"][a:b], ws["depth"][a:b], ws["hit"][a:b], ws["normal"][a:b]),
                                 queue=ws["queue"][i:i + 1])
            ev = torch.cuda.Event()
            ev.record(sc)
            s_out.wait_event(ev)
            with torch.cuda.stream(s_out):
                for k in fields:
                    out[k][a:b].copy_(ws[k][a:b], non_blocking=True)
        s_out.synchronize()             # every range's results are in host memory; all streams are idle again
        return RenderBuffer(**{k: out[k] for k in fields})

    def get_min(self, net, ray_o, ray_d):
        """Min-distance variant (reference :134-218): the aabb mask is discarded, the live mask is
        recomputed from scratch every step, per-ray (min d, x at min d) are tracked and the returned
        normals are NOT normalised.  (The reference passes `minx=` to a buffer whose field is `min_x`
        and therefore raises TypeError as shipped; here the value lands in `min_x`.)"""
        return self._forward_generic(net, ray_o, ray_d, track_min=True)

    # ------------------------------------------------------------------ generic path (torch ops, device-side)
    def _forward_generic(self, net, ray_o, ray_d, track_min):
        x, t, live = ops.aabb(ray_o, ray_d)
        far = self.camera_clamp[1]
        with torch.no_grad():
            d = net(x)
            dprev = d.clone()
            if track_min:
                live = torch.ones_like(live)
                mind, minx = d.clone(), x.clone()
            flag = torch.zeros_like(live)
            for _ in range(self.num_steps):
                flag = (t.abs() < far)[:, 0]
                tests = (d.abs() > self.min_dis)[:, 0] & (((d + dprev) / 2.0).abs() > self.min_dis * 3)[:, 0] & flag
                live = tests if track_min else (live & tests)
                if not bool(live.any()):
                    break
                col = live.unsqueeze(1)
                x = torch.where(col, torch.addcmul(ray_o, ray_d, t), x)
                if track_min:
                    lower = (d < mind)[:, 0]
                    mind[lower] = d[lower]
                    minx[lower] = x[lower]
                dprev = torch.where(col, d, dprev)
                d[live] = net(x[live]) * self.step_size
                t = torch.where(col, t + d, t)
        hit = flag & ~(x.abs() > 1.0).any(dim=-1)
        normal = torch.zeros_like(x)
        g = gradient(x[hit], net, method=self.grad_method)
        normal[hit] = g if track_min else F.normalize(g, p=2, dim=-1, eps=1e-5)
        if track_min:
            return RenderBuffer(x=x, depth=t, hit=hit, normal=normal, min_x=minx)
        return RenderBuffer(x=x, depth=t, hit=hit, normal=normal)

    def sample_surface(self, n, net):
        """Random-ray surface sampler (reference :220-245).  Origins are U[-1,1]^3, i.e. inside the
        cube, so -- exactly as in the reference (aabb leaves such rays untouched and the tracer then
        reports them as hits) -- the returned points are the origins themselves."""
        device = next(net.parameters()).device
        pts = None
        with torch.no_grad():
            for it in range(1000):
                ray_o = torch.rand((n, 3), device=device) * 2.0 - 1.0
                u = np.random.rand(2, n)
                z = 1 - 2 * u[0]
                r = np.sqrt(1.0 - z * z)
                phi = 2 * np.pi * u[1]
                ray_d = torch.from_numpy(np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1)).float().to(device)
                rb = self.forward(net, ray_o, ray_d)
                got = rb.x[rb.hit]
                pts = got if pts is None else torch.cat([pts, got], dim=0)
                if pts.shape[0] >= n:
                    break
                if it == 49:
                    print("Taking an unusually long time to sample desired # of points.")
        return pts
