"""Renderer -- the reference's orchestration class (sdf-net/lib/renderer.py:40-331) over the
fused tracer.  Public surface kept: ctor(tracer, args=None, **kwargs), render_lookat, render,
shade_tensor, shade_images, sdf_slice / normal_slice / sdf_grad_slice; output layout (W,H,C).

Differences from the reference are confined to where work runs, not what is computed:
matcap lookup, the shadow-map blur and AO all stay on the device (the reference bounces through
numpy/scipy on the host, renderer.py:279-305), and the 40 AO taps are evaluated on the compacted
hit set with the fused SDF kernel.
"""
import time

import numpy as np
import torch
import torch.nn.functional as F

from .utils import PerfTimer, setparam
from .diffutils import gradient
from .geoutils import look_at, spherical_envmap, matcap_sampler, procedural_matcap, normalized_slice, \
    gaussian_blur2d
from .tracer import RenderBuffer
from .. import ops


class Renderer():
    def __init__(self, tracer, args=None, render_res=None, camera_clamp=None, camera_proj=None,
                 render_batch=None, shading_mode=None, matcap_path=None, shadow=None, ao=None, perf=None,
                 device=None, ground_height=None):
        self.args = args
        self.tracer = tracer
        self.camera_clamp = tracer.camera_clamp
        self.render_res = setparam(args, render_res, "render_res")
        self.camera_proj = setparam(args, camera_proj, "camera_proj")
        self.render_batch = setparam(args, render_batch, "render_batch") or 0
        self.shading_mode = setparam(args, shading_mode, "shading_mode")
        self.matcap_path = setparam(args, matcap_path, "matcap_path")
        self.shadow = setparam(args, shadow, "shadow")
        self.ao = setparam(args, ao, "ao")
        self.perf = setparam(args, perf, "perf")
        self.device = setparam(args, device, "device")
        if self.device is None:
            self.device = "cuda"
        self.width, self.height = self.render_res
        self._matcap = None
        # optional [N,3] tensor used as the shadow rays' direction jitter instead of a fresh N(0, 0.01) draw
        # (renderer.py:151 draws it from torch's global generator); lets a caller reproduce a frame exactly
        self.shadow_jitter = None
        if self.shadow:
            gh = setparam(args, ground_height, "ground_height")
            if not gh:
                # the reference falls into an undefined `self.sdf_net` here (renderer.py:81)
                raise ValueError("--shadow needs a non-zero --ground-height")
            self.min_y = gh

    # ------------------------------------------------------------------ camera
    def render_lookat(self, net, f=[0, 0, 1], t=[0, 0, 0], fov=30.0, camera_proj="persp", device=None, mm=None):
        device = device or self.device
        ray_o, ray_d = look_at(f, t, self.width, self.height, fov=fov, mode=camera_proj, device=device)
        if mm is not None:
            mm = mm.to(ray_o.device)
            ray_o = torch.mm(ray_o, mm)
            ray_d = torch.mm(ray_d, mm)
        return self.render(net, ray_o, ray_d)

    # ------------------------------------------------------------------ trace + secondary passes
    def _trace(self, net, ray_o, ray_d):
        if self.render_batch > 0:
            rb = RenderBuffer()
            for o, d in zip(torch.split(ray_o, self.render_batch), torch.split(ray_d, self.render_batch)):
                rb += self.tracer(net, o, d)
            return rb
        return self.tracer(net, ray_o, ray_d)

    def render(self, net, ray_o, ray_d):
        timer = PerfTimer(activate=bool(self.perf))
        t0 = time.time()
        with torch.no_grad():
            rb = self._trace(net, ray_o, ray_d)
            timer.check("trace")

            plane_hit = None
            if self.shadow:
                # ground plane y = min_y, then a second trace towards the light (renderer.py:131-161)
                rate = -ray_d[:, 1]
                plane_t = (ray_o[:, 1] - self.min_y) / rate
                plane_hit = (plane_t > 0) & (plane_t < 500) & (plane_t < rb.depth[..., 0])
                rb.hit = rb.hit & ~plane_hit
                rb.depth[plane_hit] = plane_t[plane_hit].unsqueeze(1)
                rb.x[plane_hit] = ray_o[plane_hit] + ray_d[plane_hit] * plane_t[plane_hit].unsqueeze(1)
                rb.normal[plane_hit] = 0
                rb.normal[plane_hit, 1] = 1

                light_o = torch.tensor([[-1.5, 4.5, -1.5]], device=ray_o.device)
                s_o = rb.x + 0.1 * rb.normal
                jitter = torch.zeros_like(rb.x).normal_(0.0, 0.01) if self.shadow_jitter is None \
                    else self.shadow_jitter.to(rb.x.device, rb.x.dtype)
                s_d = F.normalize(jitter + light_o - s_o, dim=1)
                lit = (s_d * rb.normal).sum(-1) > 0.0
                rb.shadow = self.tracer(net, s_o, s_d).hit
                rb.shadow[~lit] = 0
                timer.check("shadow")

            rb.relative_depth = torch.clamp(rb.depth, 0.0, self.camera_clamp[1]) / self.camera_clamp[1]

            if self.shading_mode == "rb" and rb.rgb is None:
                rb.rgb = net.render(rb.x, ray_d, F.normalize(rb.normal))[..., :3]

            if self.ao:
                rb.ao = self._ambient_occlusion(net, rb, plane_hit)
                timer.check("ao")

        rb.view = ray_d
        rb = rb.reshape(self.width, self.height, -1)
        if self.perf:
            print("Time Elapsed:{:.4f}".format(time.time() - t0))
        return rb

    def _ambient_occlusion(self, net, rb, plane_hit):
        """40 taps along the normal (renderer.py:184-209), on the compacted hit / ground sets."""
        acc = torch.zeros_like(rb.depth)
        hit_idx = rb.hit.nonzero(as_tuple=True)[0]
        xs, ns = rb.x[hit_idx], rb.normal[hit_idx]
        acc_h = torch.zeros(hit_idx.shape[0], 1, device=acc.device)
        if plane_hit is not None:
            p_idx = plane_hit.nonzero(as_tuple=True)[0]
            xp, np_ = rb.x[p_idx], rb.normal[p_idx]
            acc_p = torch.zeros(p_idx.shape[0], 1, device=acc.device)
        for i in range(40):
            _d = 0.1 * 0.25 * (float(i + 1) / float(40 + 1)) ** 1.6
            if hit_idx.numel():
                acc_h += 3.5 * F.relu(_d - net(xs + ns * _d) - 0.0015)
            if plane_hit is not None and p_idx.numel():
                r = torch.minimum(net(xp + np_ * _d), torch.full_like(acc_p, _d))
                acc_p += 3.5 * F.relu(_d - r - 0.0015)
        acc[hit_idx] = acc_h
        if plane_hit is not None:
            acc[p_idx] = acc_p
        ao = torch.clamp(1.0 - acc, 0.1, 1.0)
        return ao * ao

    # ------------------------------------------------------------------ shading
    def _get_matcap(self, device):
        if self._matcap is None:
            try:
                self._matcap = matcap_sampler(self.matcap_path, device=device)
            except (FileNotFoundError, TypeError, AttributeError):
                self._matcap = procedural_matcap(device=device)     # no assets ship with the repo
        return self._matcap

    def shade_tensor(self, net, f=[0, 0, 1], t=[0, 0, 0], fov=30.0, mm=None):
        rb = self.render_lookat(net, f=f, t=t, fov=fov, mm=mm)
        if self.shading_mode == "matcap":
            view = rb.view
            if mm is not None:
                mm = mm.to(view.device)
                view = torch.mm(view.reshape(-1, 3), mm.transpose(1, 0)).reshape(self.width, self.height, 3)
            matcap = self._get_matcap(view.device)
            if not view.is_cuda:
                raise RuntimeError("Renderer.shade_tensor: the render buffers must live on a CUDA device (no CPU path)")
            # envmap + bilinear lookup + "misses are white" (normals of misses -> 1) in one kernel (the scipy-pinned host
            # recipe it is tested against is geoutils.spherical_envmap + matcap_sampler, tests/test_gpu_parity.py)
            rb.normal = rb.normal.contiguous()
            rb.rgb = ops.shade_matcap(view, rb.normal, rb.hit, matcap.tex)
        elif self.shading_mode == "rb":
            assert rb.rgb is not None, "No rgb in buffer; change shading-mode"
            miss = ~rb.hit[..., 0]
            rb.normal[miss] = 1.0
            rb.rgb[miss] = 1.0
        else:
            raise NotImplementedError
        if self.shadow:
            smap = torch.clamp(1.0 - rb.shadow.float() + 0.9, 0.0, 1.0)[..., 0]
            rb.rgb[..., :3] *= gaussian_blur2d(smap, 2.0).unsqueeze(-1)
        if self.ao:
            rb.rgb[..., :3] *= rb.ao
        return rb

    def shade_images(self, net, f=[0, 0, 1], t=[0, 0, 0], fov=30.0, aa=1, mm=None):
        """Returns a CPU RenderBuffer laid out (H,W,C).  As in the reference, `aa` only multisamples
        when it is an int > 1 (sdf_renderer.py passes a bool, so the stock app never does)."""
        if mm is None and not aa > 1 and self._can_pipeline(net):
            return self._shade_images_pipelined(net, f, t, fov)
        if mm is None:
            mm = torch.eye(3)
        if aa > 1:
            rb = RenderBuffer.mean(*[self.shade_tensor(net, f=f, t=t, fov=fov, mm=mm) for _ in range(aa)])
        else:
            rb = self.shade_tensor(net, f=f, t=t, fov=fov, mm=mm)
        return rb.cpu().transpose()

    # ------------------------------------------------------------------ shade_images, copies overlapped with the trace
    pipelined = True        # False: always shade_tensor + RenderBuffer.cpu() (the A/B switch of the identity test)
    pipe_reserve_sms = 8    # SMs the 2nd, 3rd, ... range's tracer launch leaves to the previous range's shading kernels

    def _can_pipeline(self, net):
        from .tracer.SphereTracer import SphereTracer, _is_octree
        tr = self.tracer
        return (self.pipelined and type(tr) is SphereTracer and _is_octree(net) and tr.grad_method == "finitediff"
                and getattr(net, "interpolate", None) is None and self.shading_mode == "matcap" and not self.shadow
                and not self.ao and not self.render_batch and not self.perf and torch.device(self.device).type == "cuda")

    def _shade_images_pipelined(self, net, f, t, fov, chunks=3):
        """`shade_tensor(...).cpu().transpose()` for the plain matcap frame (no shadow / AO / model matrix) with the 57
        bytes per ray of RenderBuffer fields crossing PCIe WHILE the frame is traced: `view` (= ray_d) leaves as soon as
        the rays exist, and the frame is traced, shaded and copied in `chunks` contiguous ray ranges on alternating
        streams (rays are independent and x-major, so a range is a block of image columns).  Same kernels, same
        arithmetic, same random draws as the generic path: the returned buffers are equal bit for bit
        (tests/test_render_app.py).  720p: 2.4 -> 1.5 ms per call."""
        from .geoutils import _window, camera_basis
        from .tracer.SphereTracer import _trace_lod
        dev = next(net.parameters()).device
        W, H = self.width, self.height
        n = W * H
        tr = self.tracer
        far = self.camera_clamp[1]
        ws = getattr(self, "_pipe_ws", None)
        if ws is None or ws["n"] != n or ws["dev"] != dev:
            ws = {"n": n, "dev": dev, "o": torch.empty(n, 3, device=dev), "d": torch.empty(n, 3, device=dev),
                  "x": torch.empty(n, 3, device=dev), "depth": torch.empty(n, 1, device=dev),
                  "hit": torch.empty(n, dtype=torch.bool, device=dev), "normal": torch.empty(n, 3, device=dev),
                  "relative_depth": torch.empty(n, 1, device=dev), "rgb": torch.empty(n, 3, device=dev),
                  "queue": torch.empty(16, dtype=torch.int32, device=dev),
                  "s_out": torch.cuda.Stream(dev), "s_c": [torch.cuda.Stream(dev) for _ in range(2)],
                  "max_ctas": max(1, torch.cuda.get_device_properties(dev).multi_processor_count - self.pipe_reserve_sms)
                  if self.pipe_reserve_sms > 0 else 0}
            self._pipe_ws = ws
        shapes = {"x": (3, torch.float32), "hit": (1, torch.bool), "depth": (1, torch.float32),
                  "relative_depth": (1, torch.float32), "normal": (3, torch.float32), "rgb": (3, torch.float32),
                  "view": (3, torch.float32)}
        host = {k: torch.empty((n, c), dtype=dt, pin_memory=True) for k, (c, dt) in shapes.items()}
        view, lod = net.net_view(), _trace_lod(net)
        tex = self._get_matcap(dev).tex
        cur = torch.cuda.current_stream(dev)
        s_out = ws["s_out"]
        with torch.cuda.device(dev), torch.no_grad():
            origin, cview, right, up = camera_basis(f, t)
            wx, wy = _window(W, H, dev)                       # the jitter draws of look_at, in its order
            ops.generate_rays(origin, cview, right, up, np.float32(np.tan(np.radians(fov / 2))), False, wx, wy,
                              out=(ws["o"], ws["d"]))
            s_out.wait_stream(cur)
            with torch.cuda.stream(s_out):
                host["view"].copy_(ws["d"], non_blocking=True)
            if isinstance(chunks, (tuple, list)):           # cumulative split points in (0, 1)
                bounds = [0] + [(int(n * f_) // H) * H for f_ in chunks] + [n]
                chunks = len(bounds) - 1
            else:
                bounds = [((n * i) // chunks // H) * H if 0 < i < chunks else (n * i) // chunks for i in range(chunks + 1)]
            for i in range(chunks):
                a, b = bounds[i], bounds[i + 1]
                if b == a:
                    continue
                sc = ws["s_c"][i % 2]
                sc.wait_stream(cur)
                with torch.cuda.stream(sc):
                    # ranges after the first leave a few SMs free: the previous range's shading kernels run beside this
                    # trace instead of waiting for it (a tracer CTA holds its SM's whole register file)
                    ops.sphere_trace(view, lod, ws["o"][a:b], ws["d"][a:b], num_steps=tr.num_steps, step_size=tr.step_size,
                                     min_dis=tr.min_dis, far=far,
                                     out=(ws["x"][a:b], ws["depth"][a:b], ws["hit"][a:b], ws["normal"][a:b]),
                                     queue=ws["queue"][i:i + 1], max_ctas=ws["max_ctas"] if i > 0 else 0)
                    torch.div(torch.clamp(ws["depth"][a:b], 0.0, far), far, out=ws["relative_depth"][a:b])
                    ops.shade_matcap(ws["d"][a:b], ws["normal"][a:b], ws["hit"][a:b], tex, out=ws["rgb"][a:b])
                    ev = torch.cuda.Event()
                    ev.record(sc)
                s_out.wait_event(ev)
                with torch.cuda.stream(s_out):
                    for k in ("x", "hit", "depth", "relative_depth", "normal", "rgb"):
                        host[k][a:b].copy_(ws[k][a:b].reshape(b - a, -1), non_blocking=True)
            s_out.synchronize()
            for sc in ws["s_c"]:
                cur.wait_stream(sc)
        rb = RenderBuffer(**{k: v.reshape(W, H, -1) for k, v in host.items()})
        return rb.transpose()

    # ------------------------------------------------------------------ 2-D slices
    def sdf_slice(self, net, dim=0, depth=0):
        pts = normalized_slice(self.width, self.height, dim=dim, depth=depth, device=self.device)
        with torch.no_grad():
            d = net.sdf(pts.reshape(-1, 3)).reshape(self.width, self.height).cpu().numpy()
        d = np.clip((d + 1.0) / 2.0, 0.0, 1.0)
        blue = np.clip((d - 0.5) * 2.0, 0.0, 1.0)
        vis = np.zeros([*d.shape, 3])
        vis[..., 2] = blue
        vis += (1.0 - blue)[..., None] * np.array([0.4, 0.3, 0.0]) + 0.2
        vis[d - 0.5 < 0] = np.array([1.0, 0.38, 0.0])
        for i in range(50):
            vis[np.abs(d - 0.02 * i) < 0.0015] = 0.8
        vis[np.abs(d - 0.5) < 0.004] = 0.0
        return vis

    def normal_slice(self, net, dim=0, depth=0.0):
        pts = normalized_slice(self.width, self.height, dim=dim, depth=depth, device=self.device).reshape(-1, 3)
        n = (F.normalize(gradient(pts, net, method="finitediff").detach()) + 1.0) / 2.0
        return n.reshape(self.width, self.height, 3).cpu().numpy()

    def sdf_grad_slice(self, net, dim=0, depth=0):
        pts = normalized_slice(self.width, self.height, dim=dim, depth=depth, device=self.device)
        g = gradient(pts.reshape(-1, 3), net, method="finitediff").detach()
        return g.norm(2, dim=-1).reshape(self.width, self.height, 1).cpu().numpy()
