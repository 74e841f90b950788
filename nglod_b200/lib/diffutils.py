"""Differential operators on an SDF callable (reference: sdf-net/lib/diffutils.py:29-84)."""
import torch

from .. import ops

_H = 1.0 / (64.0 * 3.0)


def gradient(x, f, method="autodiff"):
    """d f / d x.  'finitediff' (the default of --grad-method) is six evaluations at +-h e_k,
    h = 1/192, divided by 2h; for an OctreeSDF it is one fused kernel."""
    if method == "finitediff":
        if hasattr(f, "net_view") and getattr(f, "interpolate", None) is None:
            lod = getattr(f, "lod", None)
            lod = f.num_lods - 1 if (lod is None or not 0 <= lod < f.num_lods) else lod
            with torch.no_grad():
                return ops.sdf_finitediff(f.net_view(), lod, x.reshape(-1, 3), _H).reshape(x.shape)
        cols = []
        for k in range(3):
            e = torch.zeros(3, device=x.device)
            e[k] = _H
            cols.append(f(x + e) - f(x - e))
        return torch.cat(cols, dim=-1) / (_H * 2.0)
    if method == "autodiff":
        with torch.enable_grad():
            x = x.requires_grad_(True)
            y = f(x)
            # the fused backward is once-differentiable, so no create_graph (the reference asks for it
            # but never differentiates the normals on this path)
            return torch.autograd.grad(y, x, grad_outputs=torch.ones_like(y))[0]
    if method == "tetrahedron":
        h = _H
        ks = torch.tensor([[1.0, -1.0, -1.0], [-1.0, -1.0, 1.0], [-1.0, 1.0, -1.0], [1.0, 1.0, 1.0]], device=x.device)
        acc = 0
        for k in ks:
            acc = acc + k * f((x + k * h).detach())
        return acc / (h * 4.0)
    if method == "multilayer":
        grads = []
        with torch.enable_grad():
            x = x.requires_grad_(True)
            for y in f.sdf(x, return_lst=True):
                grads.append(torch.autograd.grad(y, x, grad_outputs=torch.ones_like(y))[0])
        return grads
    raise NotImplementedError(method)
