"""Shading kernel vs the scipy-pinned host recipe; the sdf_renderer.py entry point end to end."""
import os

import numpy as np
import pytest
import torch

from helpers import fit3_model, make_args

pytestmark = pytest.mark.gpu


def test_shade_matcap_kernel_equals_host_recipe():
    """nglod_shade_matcap (csrc/render.cu) == spherical_envmap + bilinear matcap lookup + 'misses are white' as the CPU
    branch of Renderer.shade_tensor computes them (lib/geoutils.py; that host recipe is pinned to scipy's
    RegularGridInterpolator / the reference's formulas in tests/test_host_logic.py::test_matcap_and_blur_match_scipy)."""
    from nglod_b200 import ops
    from nglod_b200.lib.geoutils import spherical_envmap, procedural_matcap, MatcapSampler
    g = torch.Generator().manual_seed(3)
    n = 200003
    view = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=1)
    normal = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=1)
    normal[:7] = torch.tensor([[0.0, 0.0, 1.0], [0.0, 1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, -1.0], [0.0, -1.0, 0.0],
                               [-1.0, 0.0, 0.0], [0.0, 0.0, 0.0]])
    hit = torch.rand(n, generator=g) < 0.7
    for tex in (procedural_matcap(64, device="cpu").tex, torch.rand(37, 53, 4, generator=g) * 255.0):
        sampler = MatcapSampler(tex)
        uv = spherical_envmap(view.clone(), normal.clone())
        ref_rgb = sampler(uv)[..., :3] / 255.0
        ref_rgb[~hit] = 1.0
        ref_n = normal.clone()
        ref_n[~hit] = 1.0
        dn = normal.clone().cuda().contiguous()
        rgb = ops.shade_matcap(view.cuda(), dn, hit.cuda(), tex.cuda())
        ok = torch.isfinite(ref_rgb).all(dim=1)                       # degenerate inputs (zero normal) may be NaN in both
        # rgb in [0,1]: 1e-4 is 1/40 of an 8-bit level (envmap: normalise + asin-free uv; bilinear weights in fp32 both sides)
        assert (rgb.cpu() - ref_rgb)[ok].abs().max() < 1e-4
        assert torch.equal(dn.cpu()[~hit], ref_n[~hit]) and torch.equal(dn.cpu()[hit], ref_n[hit])
        assert (rgb.cpu()[~hit] == 1.0).all()


def test_sdf_renderer_main_writes_the_frames_shade_images_returns(fit3, tmp_path):
    """app/sdf_renderer.py (the reference's CLI, app/sdf_renderer.py:55-204, single-frame branch): load a checkpoint by
    --pretrained, build --net / --tracer by name, set --lod, render, write <name>_{rgb,depth,normal,hit}.png -- and the
    PNGs are exactly RenderBuffer.image() of Renderer.shade_images under the same seed."""
    from PIL import Image
    from nglod_b200.app import sdf_renderer
    from nglod_b200.lib.renderer import Renderer
    from nglod_b200.lib.tracer import SphereTracer
    net, _ = fit3_model(fit3)
    ckpt = tmp_path / "torus3.pth"
    torch.save(net.state_dict(), ckpt)
    argv = ["--net", "OctreeSDF", "--num-lods", "3", "--lod", "2", "--pretrained", str(ckpt), "--render-res", "160", "90",
            "--shading-mode", "matcap", "--ao", "--img-dir", str(tmp_path / "imgs")]
    torch.manual_seed(42)
    out_dir = sdf_renderer.main(argv)
    files = sorted(os.listdir(out_dir))
    assert files == ["torus3_depth.png", "torus3_hit.png", "torus3_normal.png", "torus3_rgb.png"]
    # the same frame through the library
    args = make_args(argv[2:-2])
    dnet, _ = fit3_model(fit3, "cuda")
    dnet.lod = 2
    dnet.eval()
    r = Renderer(SphereTracer(args), args=args, device=torch.device("cuda"))
    torch.manual_seed(42)
    out = r.shade_images(net=dnet, f=args.camera_origin, t=args.camera_lookat, fov=args.camera_fov, aa=True, mm=torch.eye(3))
    img = out.image().byte().numpy()
    for key, fname in (("rgb", "torus3_rgb.png"), ("depth", "torus3_depth.png"), ("normal", "torus3_normal.png")):
        got = np.array(Image.open(os.path.join(out_dir, fname)))
        assert got.shape == (90, 160, 3)
        assert np.array_equal(got, getattr(img, key)), key
    hit_png = np.array(Image.open(os.path.join(out_dir, "torus3_hit.png")))
    assert np.array_equal(hit_png, img.hit[..., 0])
    assert 0.05 < (hit_png > 0).mean() < 0.6                            # the torus covers part of the frame
    # flags that must fail loudly
    with pytest.raises(SystemExit):
        sdf_renderer.main(["--net", "OctreeSDF", "--img-dir", str(tmp_path)])      # no --pretrained


def test_shade_images_pipelined_equals_generic_path(fit3):
    """Renderer.shade_images for the plain matcap frame overlaps the device->host copies with the trace
    (_shade_images_pipelined: chunks on alternating streams); its CPU RenderBuffer equals, field for field and bit for bit,
    `shade_tensor(...).cpu().transpose()` under the same seed (renderer.py:310-331), and options it does not cover
    (shadow / AO / aa / a model matrix) still take the generic path."""
    from nglod_b200.lib.renderer import Renderer
    from nglod_b200.lib.tracer import SphereTracer
    net, _ = fit3_model(fit3, "cuda")
    net.lod = 2
    args = make_args(["--num-lods", "3", "--lod", "2", "--render-res", "200", "113", "--shading-mode", "matcap"])
    r = Renderer(SphereTracer(args), args=args, device="cuda")
    assert r._can_pipeline(net)
    cam = dict(f=[-2.8, 2.8, -2.8], t=[0.0, 0.0, 0.0], fov=30.0)
    outs = []
    for pipelined in (True, False, True):
        r.pipelined = pipelined
        torch.manual_seed(9)
        outs.append(r.shade_images(net, **cam))
    names = [k for k in outs[1]._names() if getattr(outs[1], k) is not None]
    assert sorted(names) == ["depth", "hit", "normal", "relative_depth", "rgb", "view", "x"]
    assert int(outs[1].hit.sum()) > 800
    for other in (outs[0], outs[2]):
        assert sorted(k for k in other._names() if getattr(other, k) is not None) == sorted(names)
        for k in names:
            a, b = getattr(other, k), getattr(outs[1], k)
            assert not a.is_cuda and a.shape == b.shape == (113, 200, a.shape[-1]) and a.dtype == b.dtype, k
            assert torch.equal(a, b), k
    r.pipelined = True
    r2 = Renderer(SphereTracer(args), args=make_args(["--num-lods", "3", "--lod", "2", "--render-res", "200", "113",
                                                      "--shading-mode", "matcap", "--ao"]), device="cuda")
    assert not r2._can_pipeline(net)
