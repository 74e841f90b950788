"""Render an OctreeSDF to PNGs -- the single-frame branch of the reference's app/sdf_renderer.py:55-204.

    python -m nglod_b200.app.sdf_renderer --net OctreeSDF --num-lods 5 --pretrained m.pth \
           --render-res 1280 720 --shading-mode matcap --lod 4 [--shadow --ground-height -0.3] [--ao]

Same flags as the reference (lib/options.py + the 'app' group).  --export / --sol / --exr / --r360 are out of scope
(SOL export is broken in the reference as shipped; pyexr / moviepy are not dependencies of the hot path).
"""
import os
import sys

import numpy as np
import torch
from PIL import Image

from ..lib.options import parse_options
from ..lib.renderer import Renderer
from ..lib.models import *  # noqa: F401,F403  (classes are resolved by name, like the reference)
from ..lib.tracer import *  # noqa: F401,F403


def main(argv=None):
    parser = parse_options(return_parser=True)
    app = parser.add_argument_group("app")
    app.add_argument("--img-dir", type=str, default="_results/render_app/imgs")
    app.add_argument("--disable-aa", action="store_true")
    app.add_argument("--rotate", type=float, default=None)
    args = parser.parse_args(argv)
    if not torch.cuda.is_available():
        raise RuntimeError("nglod_b200 renders on a CUDA device only (no CPU path)")
    device = torch.device("cuda")
    if args.pretrained is None:
        raise SystemExit("No network weights specified! (--pretrained)")
    name = os.path.basename(args.pretrained).split(".")[0]
    net = globals()[args.net](args)
    net.load_state_dict(torch.load(args.pretrained, map_location="cpu"))
    net.to(device).eval()
    print("Total number of parameters: {}".format(sum(p.numel() for p in net.parameters())))
    if args.lod is not None:
        net.lod = args.lod
    out_dir = os.path.join(args.img_dir, name)
    os.makedirs(out_dir, exist_ok=True)
    tracer = globals()[args.tracer](args)
    renderer = Renderer(tracer, args=args, device=device)
    mm = torch.eye(3)
    if args.rotate is not None:
        a = np.radians(args.rotate)
        mm = torch.tensor([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]], dtype=torch.float32)
    out = renderer.shade_images(net=net, f=args.camera_origin, t=args.camera_lookat, fov=args.camera_fov,
                                aa=not args.disable_aa, mm=mm)
    img = out.image().byte().numpy()
    Image.fromarray(img.rgb).save(os.path.join(out_dir, f"{name}_rgb.png"), mode="RGB")
    Image.fromarray(img.depth).save(os.path.join(out_dir, f"{name}_depth.png"), mode="RGB")
    Image.fromarray(img.normal).save(os.path.join(out_dir, f"{name}_normal.png"), mode="RGB")
    Image.fromarray(img.hit[..., 0]).save(os.path.join(out_dir, f"{name}_hit.png"), mode="L")
    return out_dir


if __name__ == "__main__":
    sys.exit(0 if main() else 1)
