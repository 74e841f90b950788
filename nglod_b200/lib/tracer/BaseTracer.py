"""Tracer base: constructor parameters (reference: sdf-net/lib/tracer/BaseTracer.py:26-52)."""
import math

from ..utils import setparam


class BaseTracer(object):
    def __init__(self, args=None, camera_clamp=None, step_size=None, grad_method=None, num_steps=None,
                 min_dis=None):
        self.args = args
        self.camera_clamp = setparam(args, camera_clamp, "camera_clamp")
        self.step_size = setparam(args, step_size, "step_size")
        self.grad_method = setparam(args, grad_method, "grad_method")
        self.num_steps = setparam(args, num_steps, "num_steps")
        self.min_dis = setparam(args, min_dis, "min_dis")
        self.inv_num_steps = 1.0 / self.num_steps
        self.diagonal = math.sqrt(3) * 2.0

    def __call__(self, *args, **kwargs):
        return self.forward(*args, **kwargs)

    def forward(self, net, ray_o, ray_d):
        raise NotImplementedError
