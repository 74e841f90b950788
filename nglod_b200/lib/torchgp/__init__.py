"""Mesh sampling utilities (reference: sdf-net/lib/torchgp/*), device-resident."""
from .sampling import (area_weighted_distribution, random_face, point_sample, sample_surface,  # noqa: F401
                       sample_near_surface, sample_uniform, sample_spc, per_face_normals, normalize)
from .compute_sdf import compute_sdf  # noqa: F401
from .load_obj import load_obj, write_obj  # noqa: F401
from .meshes import icosphere, torus  # noqa: F401
