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
