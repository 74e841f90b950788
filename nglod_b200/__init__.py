"""nglod_b200 -- B200-native (sm_100a) implementation of NGLOD's hot path.

Drop-in surface (mirrors the reference's `sdf-net/lib` package):
    from nglod_b200.lib.models import OctreeSDF
    from nglod_b200.lib.tracer import SphereTracer, RenderBuffer
    from nglod_b200.lib.renderer import Renderer
    from nglod_b200.lib.datasets import MeshDataset
    from nglod_b200.lib.options import parse_options
or put this directory on sys.path and keep the reference's `from lib.models import *`.
The compute lives in libnglod_b200.so (C ABI: include/nglod_b200.h).
"""
__version__ = "0.1.0"
