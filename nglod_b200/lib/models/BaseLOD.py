"""LOD-aware base (reference: sdf-net/lib/models/BaseLOD.py:31-47)."""
from .BaseSDF import BaseSDF


class BaseLOD(BaseSDF):
    def __init__(self, args):
        super().__init__(args)
        self.num_lods = args.num_lods
        self.lod = None

    def forward(self, x, lod=None):
        # as in the reference (BaseLOD.py:37-41) the `lod` argument is NOT forwarded:
        # `net(x)` always evaluates at `net.lod`.
        return self.sdf(self.encode(x))
