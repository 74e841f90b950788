"""compute_sdf (reference: sdf-net/lib/torchgp/compute_sdf.py:26-40)."""
from ... import ops


def compute_sdf(V, F, points):
    """[N,3] points -> [N] signed distances to the mesh (V, F), via the sm_100a mesh2sdf kernel."""
    mesh = V[F]
    return ops.mesh2sdf_gpu(points.contiguous(), mesh)[0]
