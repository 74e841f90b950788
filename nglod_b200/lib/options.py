"""CLI flag system -- same flags, defaults and groups as the reference
(sdf-net/lib/options.py:30-238) so existing command lines keep working.
The table below is (group, flag, kwargs); `parse_options` turns it into argparse."""
import argparse
import pprint

_S = dict(action="store_true")

_FLAGS = [
    # -------- global
    ("global", "--exp-name", dict(type=str)),
    ("global", "--perf", _S),
    ("global", "--validator", dict(type=str, default=None)),
    ("global", "--valid-only", _S),
    ("global", "--valid-every", dict(type=int, default=1)),
    ("global", "--debug", _S),
    ("global", "--seed", dict(type=int)),
    ("global", "--ngc", _S),
    # -------- net
    ("net", "--net", dict(type=str, default="OverfitSDF")),
    ("net", "--jit", _S),
    ("net", "--pos-enc", _S),
    ("net", "--feature-dim", dict(type=int, default=32)),
    ("net", "--feature-size", dict(type=int, default=4)),
    ("net", "--joint-feature", _S),
    ("net", "--num-layers", dict(type=int, default=1)),
    ("net", "--num-lods", dict(type=int, default=1)),
    ("net", "--base-lod", dict(type=int, default=2)),
    ("net", "--ff-dim", dict(type=int, default=-1)),
    ("net", "--ff-width", dict(type=float, default=16.0)),
    ("net", "--hidden-dim", dict(type=int, default=128)),
    ("net", "--pretrained", dict(type=str)),
    ("net", "--periodic", _S),
    ("net", "--skip", dict(type=int, default=None)),
    ("net", "--freeze", dict(type=int, default=-1)),
    ("net", "--pos-invariant", _S),
    ("net", "--joint-decoder", _S),
    # not in the reference: kernel knobs of this implementation (defaults keep the reference's numerics)
    ("net", "--math-mode", dict(type=str, default=None, choices=["tc", "fp32"])),
    ("net", "--grid-storage", dict(type=str, default=None, choices=["fp32", "fp16"])),
    ("net", "--no-sum-lods", dict(dest="sum_lods", action="store_false", default=None)),
    ("net", "--feat-sum", _S),
    # -------- dataset
    ("dataset", "--dataset-path", dict(type=str)),
    ("dataset", "--analytic", _S),
    ("dataset", "--mesh-dataset", dict(type=str, default="MeshDataset")),
    ("dataset", "--raw-obj-path", dict(type=str, default=None)),
    ("dataset", "--mesh-batch", _S),
    ("dataset", "--mesh-subset-size", dict(type=int, default=-1)),
    ("dataset", "--train-valid-split", dict(type=str, default=None)),
    ("dataset", "--num-samples", dict(type=int, default=100000)),
    ("dataset", "--samples-per-voxel", dict(type=int, default=256)),
    ("dataset", "--sample-mode", dict(type=str, nargs="*", default=["rand", "near", "near", "trace", "trace"])),
    ("dataset", "--trim", _S),
    ("dataset", "--sample-tex", _S),
    ("dataset", "--block-res", dict(type=int, default=7)),
    ("dataset", "--include", dict(nargs="*")),
    ("dataset", "--exclude", dict(nargs="*")),
    ("dataset", "--glsl-path", dict(type=str, default="../sdf-viewer/data-files/sdf")),
    ("dataset", "--viewer-path", dict(type=str, default="../sdf-viewer")),
    ("dataset", "--get-normals", _S),
    ("dataset", "--build-dataset", _S),
    # -------- optimizer
    ("optimizer", "--optimizer", dict(type=str, default="adam", choices=["adam", "sgd"])),
    ("optimizer", "--lr", dict(type=float, default=0.001)),
    ("optimizer", "--loss", dict(nargs="+", type=str, default=["l2_loss"])),
    ("optimizer", "--grad-method", dict(type=str, choices=["autodiff", "finitediff"], default="finitediff")),
    # -------- trainer
    ("trainer", "--epochs", dict(type=int, default=250)),
    ("trainer", "--batch-size", dict(type=int, default=512)),
    ("trainer", "--only-last", _S),
    ("trainer", "--resample-every", dict(type=int, default=10)),
    ("trainer", "--model-path", dict(type=str, default="_results/models")),
    ("trainer", "--save-as-new", _S),
    ("trainer", "--save-every", dict(type=int, default=1)),
    ("trainer", "--save-all", _S),
    ("trainer", "--latent", _S),
    ("trainer", "--return-lst", _S),
    ("trainer", "--latent-dim", dict(type=int, default=128)),
    ("trainer", "--logs", dict(type=str, default="_results/logs/runs/")),
    ("trainer", "--grow-every", dict(type=int, default=-1)),
    ("trainer", "--loss-sample", dict(type=int, default=-1)),
    ("trainer", "--growth-strategy", dict(type=str, default="increase",
                                          choices=["onebyone", "increase", "shrink", "finetocoarse", "onlylast"])),
    # -------- renderer
    ("renderer", "--sol", _S),
    ("renderer", "--render-res", dict(type=int, nargs=2, default=[512, 512])),
    ("renderer", "--render-batch", dict(type=int, default=0)),
    ("renderer", "--matcap-path", dict(type=str, default="data/matcap/green.png")),
    ("renderer", "--camera-origin", dict(type=float, nargs=3, default=[-2.8, 2.8, -2.8])),
    ("renderer", "--camera-lookat", dict(type=float, nargs=3, default=[0, 0, 0])),
    ("renderer", "--camera-fov", dict(type=float, default=30)),
    ("renderer", "--camera-proj", dict(type=str, choices=["ortho", "persp"], default="persp")),
    ("renderer", "--camera-clamp", dict(nargs=2, type=float, default=[-5, 10])),
    ("renderer", "--lod", dict(type=int, default=None)),
    ("renderer", "--interpolate", dict(type=float, default=None)),
    ("renderer", "--render-every", dict(type=int, default=1)),
    ("renderer", "--num-steps", dict(type=int, default=256)),
    ("renderer", "--step-size", dict(type=float, default=1.0)),
    ("renderer", "--min-dis", dict(type=float, default=0.0003)),
    ("renderer", "--ground-height", dict(type=float)),
    ("renderer", "--tracer", dict(type=str, default="SphereTracer")),
    ("renderer", "--ao", _S),
    ("renderer", "--shadow", _S),
    ("renderer", "--shading-mode", dict(type=str, default="matcap")),
]


def parse_options(return_parser=False):
    """Reference: options.py:30.  `return_parser=True` hands back the parser so apps can add an
    'app' group; otherwise parses sys.argv and returns (args, markdown-ish string of the args)."""
    parser = argparse.ArgumentParser(description="Train / render neural SDFs (NGLOD hot path, B200-native).")
    groups = {}
    for group, flag, kw in _FLAGS:
        if group not in groups:
            groups[group] = parser.add_argument_group(group)
        groups[group].add_argument(flag, **kw)
    return parser if return_parser else argparse_to_str(parser)


def argparse_to_str(parser, argv=None):
    """Reference: options.py:241-259 -- args grouped by argument group, pretty-printed, fenced."""
    args = parser.parse_args(argv)
    grouped = {}
    for group in parser._action_groups:
        grouped[group.title] = {a.dest: getattr(args, a.dest, None) for a in group._group_actions}
    return args, "```" + pprint.PrettyPrinter(indent=2).pformat(grouped) + "```"
