"""Model base (reference: sdf-net/lib/models/BaseSDF.py:33-73) -- API surface only."""
import torch.nn as nn

from ..utils import setparam


class BaseSDF(nn.Module):
    def __init__(self, args=None, pos_enc=None, ff_dim=None, ff_width=None):
        super().__init__()
        self.args = args
        self.pos_enc = setparam(args, pos_enc, "pos_enc")
        self.ff_dim = setparam(args, ff_dim, "ff_dim")
        self.ff_width = setparam(args, ff_width, "ff_width")
        self.input_dim = 3
        self.out_dim = 1
        if (self.ff_dim is not None and self.ff_dim > 0) or self.pos_enc:
            # BaseSDF.py:48-54 widens input_dim for these encodings, but OctreeSDF disables
            # `encode` (OctreeSDF.py:90-92) so the reference itself cannot run such a model.
            raise NotImplementedError("positional / Fourier input encodings are outside the OctreeSDF hot path")

    def forward(self, x, lod=None):
        return self.sdf(self.encode(x))

    def freeze(self):
        for p in self.parameters():
            p.requires_grad_(False)

    def encode(self, x):
        return x

    def sdf(self, x, lod=None):
        return None
