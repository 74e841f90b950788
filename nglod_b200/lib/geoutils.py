"""Camera rays and matcap lookup (reference: sdf-net/lib/geoutils.py:140-206,237-275).

Everything stays on the device: the reference's matcap path copies UVs to the
host, interpolates with scipy and copies colours back (renderer.py:279-290);
here the same bilinear interpolation (RegularGridInterpolator 'linear' on a
[0,1]^2 lattice) is a handful of tensor ops on the GPU.
"""
import numpy as np
import torch
import torch.nn.functional as F


def normalized_grid(width, height, device="cuda"):
    """grid[x, y] -> window coordinates; one uniform jitter per column and per row, drawn with
    torch.rand on `device` in the reference's order (columns first), so a seeded reference and a
    seeded run of this function on the same device type produce identical rays."""
    wx = torch.linspace(-1, 1, steps=width, device=device) * (width / height)
    wx += torch.rand(*wx.shape, device=device) * (1.0 / width)
    wy = torch.linspace(1, -1, steps=height, device=device)
    wy += torch.rand(*wy.shape, device=device) * (1.0 / height)
    gx, gy = torch.meshgrid(wx, wy, indexing="ij")
    return torch.stack([gx, gy], dim=-1)          # [W, H, 2]: ray index = ix*H + iy (x-major)


def normalized_slice(width, height, dim=0, depth=0.0, device="cuda"):
    window = normalized_grid(width, height, device)
    plane = torch.full((width, height, 1), float(depth), device=device)
    order = {0: (plane, window[..., 0:1], window[..., 1:2]),
             1: (window[..., 0:1], plane, window[..., 1:2]),
             2: (window[..., 0:1], window[..., 1:2], plane)}
    if dim not in order:
        raise ValueError("dim is invalid!")
    pts = torch.cat(order[dim], dim=-1)
    pts[..., 1] *= -1
    return pts


_WINDOW_BASE = {}


def _window(width, height, device):
    """normalized_grid's two jittered coordinate vectors (one torch.rand per column, then one per row).  The un-jittered
    coordinates depend on (width, height, device) only and are kept: 6 small launches per frame instead of 10, the same
    draws and the same roundings (base + rand * (1 / n))."""
    key = (int(width), int(height), str(torch.device(device)))
    base = _WINDOW_BASE.get(key)
    if base is None:
        if len(_WINDOW_BASE) > 16:
            _WINDOW_BASE.clear()
        base = (torch.linspace(-1, 1, steps=width, device=device) * (width / height),
                torch.linspace(1, -1, steps=height, device=device))
        _WINDOW_BASE[key] = base
    wx = base[0] + torch.rand(width, device=device) * (1.0 / width)
    wy = base[1] + torch.rand(height, device=device) * (1.0 / height)
    return wx, wy


def camera_basis(f, t):
    """(origin, view, right, up) of look_at (reference :180-188) as four lists of Python floats, computed on the host by
    the library (nglod_camera_basis: the float32 arithmetic of torch's CPU kernels, equal bit for bit to
    `F.normalize(torch.linalg.cross(...))` on the host -- tests/test_host_logic.py) in ~3 us instead of eight torch calls."""
    import ctypes
    from .. import _lib
    lib = _lib.load()
    a, b, out = (ctypes.c_float * 3)(*f), (ctypes.c_float * 3)(*t), (ctypes.c_float * 12)()
    _lib.check(lib.nglod_camera_basis(a, b, out), "nglod_camera_basis")
    v = list(out)
    return v[0:3], v[3:6], v[6:9], v[9:12]


def look_at(f, t, width, height, mode="ortho", fov=90.0, device="cuda"):
    """Ray origins / directions [W*H, 3] for a camera at `f` looking at `t` (reference :180-206).
    On a CUDA device the [W,H] expansion is one kernel (nglod_generate_rays); the jitter is still drawn with
    torch.rand on the device in the reference's order."""
    if mode not in ("ortho", "persp"):
        raise ValueError("Invalid camera mode!")
    dev = torch.device(device)
    if dev.type == "cuda":
        from .. import ops
        origin, view, right, up = camera_basis(f, t)
        with torch.cuda.device(dev):
            wx, wy = _window(width, height, dev)
        tan = np.float32(np.tan(np.radians(fov / 2)))
        return ops.generate_rays(origin, view, right, up, tan, mode == "ortho", wx, wy)
    origin = torch.tensor(list(f), dtype=torch.float32, device=device)
    view = F.normalize(torch.tensor(list(t), dtype=torch.float32, device=device) - origin, dim=0)
    world_up = torch.tensor([0.0, 1.0, 0.0], device=device)
    right = F.normalize(torch.linalg.cross(view, world_up), dim=0)
    up = F.normalize(torch.linalg.cross(right, view), dim=0)

    coord = normalized_grid(width, height, device=device)
    tan = np.tan(np.radians(fov / 2))
    plane = right * coord[..., 0, None] * tan + up * coord[..., 1, None] * tan + origin + view
    plane = plane.reshape(-1, 3)
    if mode == "ortho":
        ray_d = F.normalize(view.unsqueeze(0).repeat(plane.shape[0], 1), dim=-1)
        ray_o = plane
    else:
        ray_d = F.normalize(plane - origin, dim=-1)
        ray_o = origin.repeat(ray_d.shape[0], 1)
    return ray_o, ray_d


def spherical_envmap(ray_dir, normal):
    """Matcap UVs from the reflected view direction (reference :253-275). [...,3] -> [N,2]."""
    d = ray_dir.clone()
    d[..., 2] *= -1
    r = d - 2.0 * torch.sum(normal * d, dim=-1, keepdim=True) * normal
    r[..., 2] -= 1.0
    m = 2.0 * torch.sqrt(torch.sum(r ** 2, dim=-1, keepdim=True))
    uv = 1.0 - ((r[..., :2] / m) + 0.5)
    uv = torch.clamp(uv.reshape(-1, 2), 0.0, 1.0)
    uv[torch.isnan(uv)] = 0
    return uv


class MatcapSampler:
    """Bilinear matcap lookup on the device; texture indexed [u, v] like the reference's
    `np.array(img).transpose(1,0,2)` + RegularGridInterpolator over linspace(0,1) axes."""

    def __init__(self, texture):
        self.tex = texture                     # [U, V, C] float tensor

    @classmethod
    def from_file(cls, path, device="cuda"):
        from PIL import Image
        arr = np.array(Image.open(path)).transpose(1, 0, 2)
        return cls(torch.from_numpy(arr.astype(np.float32)).to(device))

    def __call__(self, uv):
        tex = self.tex.to(uv.device)
        nu, nv = tex.shape[0], tex.shape[1]
        fu = uv[:, 0].clamp(0, 1) * (nu - 1)
        fv = uv[:, 1].clamp(0, 1) * (nv - 1)
        u0 = fu.floor().long().clamp(max=nu - 2) if nu > 1 else torch.zeros_like(fu, dtype=torch.long)
        v0 = fv.floor().long().clamp(max=nv - 2) if nv > 1 else torch.zeros_like(fv, dtype=torch.long)
        a = (fu - u0).unsqueeze(1)
        b = (fv - v0).unsqueeze(1)
        u1 = (u0 + 1).clamp(max=nu - 1)
        v1 = (v0 + 1).clamp(max=nv - 1)
        return (tex[u0, v0] * (1 - a) * (1 - b) + tex[u1, v0] * a * (1 - b)
                + tex[u0, v1] * (1 - a) * b + tex[u1, v1] * a * b)


def matcap_sampler(path, interpolate=True, device="cuda"):
    return MatcapSampler.from_file(path, device=device)


def procedural_matcap(size=256, device="cuda"):
    """A lit-sphere texture generated in-run (no assets ship with either repo)."""
    ax = torch.linspace(-1, 1, size, device=device)
    u, v = torch.meshgrid(ax, ax, indexing="ij")
    r2 = (u * u + v * v).clamp(max=1.0)
    nz = torch.sqrt(1.0 - r2)
    light = F.normalize(torch.tensor([-0.4, -0.5, 0.75], device=device), dim=0)
    diff = (u * light[0] + v * light[1] + nz * light[2]).clamp(min=0.0)
    spec = diff ** 24
    base = torch.tensor([70.0, 160.0, 90.0], device=device)
    rgb = (base * (0.25 + 0.75 * diff.unsqueeze(-1)) + 255.0 * 0.6 * spec.unsqueeze(-1)).clamp(0, 255)
    return MatcapSampler(rgb)


def _pad_symmetric(x, radius, dim):
    """scipy's 'reflect' boundary (= numpy 'symmetric': edge sample repeated), any radius."""
    n = x.shape[dim]
    idx = torch.arange(-radius, n + radius, device=x.device)
    period = 2 * n
    idx = idx % period
    idx = torch.where(idx >= n, period - 1 - idx, idx)
    return x.index_select(dim, idx)


def gaussian_blur2d(img, sigma):
    """scipy.ndimage.gaussian_filter(img, sigma) (mode='reflect', truncate=4.0) on the device."""
    radius = int(4.0 * sigma + 0.5)
    k = torch.arange(-radius, radius + 1, device=img.device, dtype=img.dtype)
    w = torch.exp(-0.5 * (k / sigma) ** 2)
    w = w / w.sum()
    x = _pad_symmetric(img, radius, 0)[None, None]
    x = F.conv2d(x, w.view(1, 1, -1, 1))[0, 0]
    x = _pad_symmetric(x, radius, 1)[None, None]
    x = F.conv2d(x, w.view(1, 1, 1, -1))[0, 0]
    return x


def sample_unif_sphere(n):
    u = np.random.rand(2, n)
    z = 1 - 2 * u[0, :]
    r = np.sqrt(1.0 - z * z)
    phi = 2 * np.pi * u[1, :]
    return np.array([r * np.cos(phi), r * np.sin(phi), z]).transpose()


def sample_fib_sphere(n):
    i = np.arange(0, n, dtype=float) + 0.5
    phi = np.arccos(1 - 2 * i / n)
    theta = 2.0 * np.pi * i / ((1 + 5 ** 0.5) / 2)
    return np.array([np.cos(theta) * np.sin(phi), np.sin(theta) * np.sin(phi), np.cos(phi)]).transpose()

