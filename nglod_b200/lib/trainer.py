"""Training step for OctreeSDF on the fused kernels (reference: Trainer.step_geometry, sdf-net/lib/trainer.py:294-340,
optimizer set-up :178-189).  Only the step semantics are reproduced here -- the epoch loop, TensorBoard and
checkpoint naming are the caller's business.

FusedTrainer keeps ALL parameters in one flat fp32 buffer (each nn.Parameter is a view into it; grids stay
channels-last), with a matching flat gradient buffer and flat Adam moments:
  * step = zero the flat gradient, one fused forward+loss+backward launch per loss LOD writing straight into
    the gradient views (no autograd graph, no saved activations), ONE all-reduce over the flat buffer when
    data-parallel, one Adam kernel over the flat buffer;
  * loss = sum_{lod in loss_lods} sum_i (sdf_lod(x_i) - gt_i)^2 / global_batch        (trainer.py:325-336).
"""
import torch

from .. import ops
from .. import dist as ndist


class FusedTrainer:
    def __init__(self, net, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, loss_lods=None):
        self.net = net
        self.lr, self.betas, self.eps = lr, betas, eps
        self.loss_lods = list(range(net.num_lods)) if loss_lods is None else list(loss_lods)
        params = list(net.parameters())
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedTrainer needs the model on a CUDA device (no CPU path)")
        sizes = [((p.numel() + 3) // 4) * 4 for p in params]          # keep every view 16-byte aligned
        total = sum(sizes)
        self.flat = torch.zeros(total, device=dev)
        self.flat_grad = torch.zeros(total, device=dev)
        self.exp_avg = torch.zeros(total, device=dev)
        self.exp_avg_sq = torch.zeros(total, device=dev)
        self.step_count = 0
        off = 0
        self._grad_views = {}
        for p, sz in zip(params, sizes):
            n = p.numel()
            if p.dim() == 5:      # feature grid: logical [1,F,D,H,W], physical channels-last [D,H,W,F]
                _, f, d, h, w = p.shape
                view = self.flat[off:off + n].view(1, d, h, w, f).permute(0, 4, 1, 2, 3)
                gview = self.flat_grad[off:off + n].view(1, d, h, w, f).permute(0, 4, 1, 2, 3)
            else:
                view = self.flat[off:off + n].view(p.shape)
                gview = self.flat_grad[off:off + n].view(p.shape)
            view.copy_(p.data)
            p.data = view
            p.grad = gview
            self._grad_views[p] = gview
            off += sz
        self.loss = torch.zeros(1, device=dev)

    def _grad_lists(self):
        net = self.net
        grid_grads = [self._grad_views[f.fm] for f in net.features]
        dec_grads = [tuple(self._grad_views[p] for p in net.decoder_params(i)) for i in range(net.num_lods)]
        return grid_grads, dec_grads

    def step(self, pts, gts, global_batch=None):
        """One optimisation step on this rank's slice (pts [B,3], gts [B,1] on the device).
        Returns the device scalar holding this rank's share of the loss (already divided by global_batch)."""
        net = self.net
        batch = pts.shape[0] if global_batch is None else global_batch
        self.flat_grad.zero_()
        self.loss.zero_()
        mask = 0
        for l in self.loss_lods:
            mask |= 1 << l
        grid_grads, dec_grads = self._grad_lists()
        ops.sdf_train_step(net.net_view(), mask, pts, gts, 1.0 / batch, grid_grads, dec_grads, self.loss)
        ndist.allreduce_sum_(self.flat_grad)
        self.step_count += 1
        ops.adam_step(self.flat, self.flat_grad, self.exp_avg, self.exp_avg_sq, self.step_count, lr=self.lr,
                      beta1=self.betas[0], beta2=self.betas[1], eps=self.eps)
        return self.loss
