"""Drop-in for the reference's `sol_nglod` extension module (sdf-net/lib/extensions/sol_nglod).
Put nglod_b200/shims on sys.path and the reference's `from sol_nglod import aabb` (SphereTracer.py:37)
binds to the sm_100a kernel."""
from nglod_b200.ops import aabb  # noqa: F401
