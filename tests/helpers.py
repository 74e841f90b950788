"""Shared test helpers: build models the way the golden generator did, convert between layouts."""
import numpy as np
import torch

from nglod_b200.lib.options import parse_options
from nglod_b200.lib.models import OctreeSDF


def make_args(extra=()):
    return parse_options(return_parser=True).parse_args(["--net", "OctreeSDF", "--feature-dim", "32", *extra])


def rand5_model(device="cpu"):
    """The random-init 5-LOD model of tests/golden/rand5.npz: same seed, same init order as the reference."""
    args = make_args(["--num-lods", "5"])
    torch.manual_seed(0)
    net = OctreeSDF(args)
    return net.to(device), args


def fit3_model(fit3, device="cpu"):
    args = make_args(["--num-lods", "3"])
    net = OctreeSDF(args)
    sd = {k[3:]: torch.from_numpy(v) for k, v in fit3.items() if k.startswith("sd.")}
    net.load_state_dict(sd)
    return net.to(device), args


def fit5_model(fit5, device="cpu"):
    """BASELINE config 2's shape (5 LODs, feature-dim 32) with the fitted, fp16-exact weights of tests/golden/fit5.npz."""
    args = make_args(["--num-lods", "5"])
    net = OctreeSDF(args)
    sd = {k[3:]: torch.from_numpy(v.astype(np.float32)) for k, v in fit5.items() if k.startswith("sd.")}
    net.load_state_dict(sd)
    return net.to(device), args


def weights_checksum(net):
    ps = list(net.parameters())
    return np.array([float(sum(p.detach().double().sum() for p in ps)),
                     float(sum((p.detach().double() ** 2).sum() for p in ps))])


def cl_flat(g):
    """[1,C,D,H,W] gradient -> channels-last flat vector (the kernels' physical order)."""
    return g[0].permute(1, 2, 3, 0).contiguous().reshape(-1)


def torus_sdf(p, major=0.6, minor=0.25):
    q = torch.sqrt(p[..., 0] ** 2 + p[..., 2] ** 2) - major
    return torch.sqrt(q * q + p[..., 1] ** 2) - minor
