"""Drop-in for the reference's `mesh2sdf` extension module (sdf-net/lib/extensions/mesh2sdf_cuda).
`mesh2sdf.mesh2sdf_gpu(points, mesh)` (compute_sdf.py:38) -> [dist]."""
from nglod_b200.ops import mesh2sdf_gpu  # noqa: F401
