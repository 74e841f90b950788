"""Training step for OctreeSDF on the fused kernels (reference: Trainer.step_geometry, sdf-net/lib/trainer.py:294-340,
optimizer set-up :178-189).  Only the step semantics are reproduced here -- the epoch loop, TensorBoard and
checkpoint naming are the caller's business.

FusedTrainer keeps ALL parameters in one flat fp32 buffer (each nn.Parameter is a view into it; grids stay
channels-last), with a matching flat gradient buffer and flat Adam moments:
  * step = zero the flat gradient, one fused forward+loss+backward launch per loss LOD writing straight into
    the gradient views (no autograd graph, no saved activations), then -- data-parallel -- the optimiser is SHARDED:
    reduce-scatter of the flat gradient (each rank receives the sum of its 1/N slice), the Adam kernel over that
    slice only (moments exist only for it), all-gather of the updated parameters.  Same bytes over NVLink as one
    all-reduce, 1/N of the optimiser's HBM traffic and state per rank; `shard_optimizer=False` keeps the replicated
    all-reduce + full Adam;
  * loss = sum_{lod in loss_lods} sum_i (sdf_lod(x_i) - gt_i)^2 / global_batch        (trainer.py:325-336);
  * small batches (the reference trains with --batch-size 512) are launch-bound: ~20 launches and ~0.4 ms of Python per
    step against ~0.2 ms of GPU work.  Everything up to the all-reduce is therefore captured ONCE per (batch size, loss
    LODs) in a CUDA graph and replayed (inputs copied into static buffers); batches under `summed_min_batch` points skip
    the prefix-summed grids (their per-step rebuild + restriction cascade cost more than five short gathers).
"""
import torch

from .. import ops, _lib
from .. import dist as ndist


class FusedTrainer:
    def __init__(self, net, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, loss_lods=None, use_graph=True,
                 graph_max_batch=131072, summed_min_batch=32768, shard_optimizer=True):
        if getattr(net, "padded", False):
            raise RuntimeError("FusedTrainer writes gradients straight into the parameters' flat buffer: the model must have the "
                               "kernels' own shape (feature-dim 32, hidden-dim 128); smaller models train through autograd "
                               "(Trainer falls back to torch.optim.Adam)")
        self.net = net
        self.use_graph, self.graph_max_batch, self.summed_min_batch = use_graph, graph_max_batch, summed_min_batch
        self._graphs = {}
        self.lr, self.betas, self.eps = lr, betas, eps
        self.loss_lods = list(range(net.num_lods)) if loss_lods is None else list(loss_lods)
        params = list(net.parameters())
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedTrainer needs the model on a CUDA device (no CPU path)")
        sizes = [((p.numel() + 3) // 4) * 4 for p in params]          # keep every view 16-byte aligned
        total = sum(sizes)
        self.world = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1
        self.rank = torch.distributed.get_rank() if self.world > 1 else 0
        self.sharded = bool(shard_optimizer) and self.world > 1
        if self.sharded:                                               # equal 16-byte aligned slices
            total = ((total + 4 * self.world - 1) // (4 * self.world)) * (4 * self.world)
        self.flat = torch.zeros(total, device=dev)
        self.flat_grad = torch.zeros(total, device=dev)
        self.shard_len = total // self.world if self.sharded else total
        self.shard_off = self.rank * self.shard_len if self.sharded else 0
        self.grad_shard = torch.zeros(self.shard_len, device=dev) if self.sharded else self.flat_grad
        self.exp_avg = torch.zeros(self.shard_len, device=dev)          # moments only for the slice this rank updates
        self.exp_avg_sq = torch.zeros(self.shard_len, device=dev)
        self.step_count = 0
        self.comm_ms = None                                            # set by step(time_comm=True)
        off = 0
        self._grad_views = {}
        self._param_ranges = []                                        # (parameter, offset, padded length) in the flat buffers
        for p, sz in zip(params, sizes):
            self._param_ranges.append((p, off, sz))
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
        self.lod_loss = torch.zeros(net.num_lods, device=dev)

    def _grad_lists(self):
        net = self.net
        grid_grads = [self._grad_views[f.fm] for f in net.features]
        dec_grads = [tuple(self._grad_views[p] for p in net.decoder_params(i)) for i in range(net.num_lods)]
        return grid_grads, dec_grads

    def _compute_grads(self, pts, gts, batch, lods):
        """Zero the flat gradient, run the fused forward+loss+backward of every loss LOD into it, sum the per-LOD losses."""
        net = self.net
        self.flat_grad.zero_()
        self.lod_loss.zero_()
        grid_grads, dec_grads = self._grad_lists()
        # big batches: rebuild the prefix-summed grids from this step's weights (~0.1 ms) and scatter into ONE grid
        view = net.net_view(inference=False, use_summed=pts.shape[0] >= self.summed_min_batch)
        # the gradient buffers were just zeroed: they double as the single-grid path's scratch (summed[i] aliases grids[i],
        # no copy-out pass); a frozen grid (no gradient view) keeps the model's own zero-filled scratch
        scratch = None
        if view.summed is not None:
            own = net.summed_grad_scratch()
            scratch = [g if g is not None else own[i] for i, g in enumerate(grid_grads)]
        # one call: a fused forward+loss+backward launch per LOD head (each with its own loss cell), then ONE restriction
        # cascade for all of them
        mask = 0
        for l in lods:
            mask |= 1 << l
        ops.sdf_train_step(view, mask | _lib.LOSS_PER_LOD, pts, gts, 1.0 / batch, grid_grads, dec_grads, self.lod_loss,
                           summed_scratch=scratch, scatter_scratch=net.scatter_scratch() if scratch is not None else None)
        torch.sum(self.lod_loss, dim=0, keepdim=True, out=self.loss)

    def _signature(self):
        d = self.net._derived
        return (self.flat.data_ptr(), self.flat_grad.data_ptr(), tuple(t.data_ptr() for t in d[1]) if d is not None else ())

    def _capture(self, key, pts, gts, batch, lods):
        """Capture _compute_grads for this batch shape in a CUDA graph (static input buffers).  Returns False if capture
        is not possible here; the step then simply runs eagerly."""
        try:
            st = {"pts": pts.detach().clone().contiguous(), "gts": gts.detach().clone().contiguous()}
            cur = torch.cuda.current_stream(pts.device)
            side = torch.cuda.Stream(pts.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):                  # warm-up off the capturing stream: every buffer gets allocated
                for _ in range(2):
                    self.net.mark_grids_dirty()
                    self._compute_grads(st["pts"], st["gts"], batch, lods)
            cur.wait_stream(side)
            self.net.mark_grids_dirty()                    # so the captured work includes the summed-grid rebuild
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._compute_grads(st["pts"], st["gts"], batch, lods)
            st["graph"], st["sig"] = graph, self._signature()
            return st
        except Exception:   # noqa: BLE001 -- capture is an optimisation, never a requirement
            torch.cuda.synchronize(pts.device)
            return False

    def _trainable_ranges(self):
        """[start, end) runs of the flat buffers that belong to parameters with requires_grad=True (merged).  Frozen
        parameters (`net.freeze()`, `--freeze`, trainer.py:247-251) get no Adam update -- torch.optim skips parameters
        without a gradient; with NOTHING left to train the step raises like `loss.backward()` does in the reference."""
        key = tuple(p.requires_grad for p, _, _ in self._param_ranges)
        if getattr(self, "_ranges_key", None) != key:
            runs = []
            for p, off, sz in self._param_ranges:
                if p.requires_grad:
                    if runs and runs[-1][1] == off:
                        runs[-1][1] = off + sz
                    else:
                        runs.append([off, off + sz])
            self._ranges_key, self._ranges = key, [tuple(r) for r in runs]
        if not self._ranges:
            raise RuntimeError("FusedTrainer.step: no parameter requires grad (the network is frozen)")
        return self._ranges

    def _adam(self, lo, hi, base):
        """Adam over flat[lo:hi); `base` = flat offset of element 0 of the gradient / moment buffers in use."""
        ops.adam_step(self.flat[lo:hi], self.grad_shard[lo - base:hi - base], self.exp_avg[lo - base:hi - base],
                      self.exp_avg_sq[lo - base:hi - base], self.step_count, lr=self.lr, beta1=self.betas[0],
                      beta2=self.betas[1], eps=self.eps)

    def _reduce_and_update(self):
        """Gradient exchange + Adam.  Replicated: all-reduce(flat gradient), Adam over everything.  Sharded: reduce-scatter,
        Adam over this rank's slice, all-gather of the parameters."""
        dist = torch.distributed
        runs = self._trainable_ranges()
        self.step_count += 1
        if self.sharded:
            dist.reduce_scatter_tensor(self.grad_shard, self.flat_grad, op=dist.ReduceOp.SUM)
            s0, s1 = self.shard_off, self.shard_off + self.shard_len
            for lo, hi in runs:
                lo, hi = max(lo, s0), min(hi, s1)
                if hi > lo:
                    self._adam(lo, hi, s0)
            dist.all_gather_into_tensor(self.flat, self.flat[s0:s1].clone())
        else:
            ndist.allreduce_sum_(self.flat_grad)
            for lo, hi in runs:
                self._adam(lo, hi, 0)

    def step(self, pts, gts, global_batch=None, loss_lods=None):
        """One optimisation step on this rank's slice (pts [B,3], gts [B,1] on the device).
        Returns the device scalar holding this rank's share of the loss (already divided by global_batch);
        `self.lod_loss[l]` holds LOD l's share."""
        net = self.net
        batch = pts.shape[0] if global_batch is None else global_batch
        lods = self.loss_lods if loss_lods is None else list(loss_lods)
        g = None
        if self.use_graph and pts.shape[0] <= self.graph_max_batch and pts.is_cuda and pts.dtype == torch.float32:
            key = (pts.shape[0], int(batch), tuple(lods), bool(net.sum_lods))
            g = self._graphs.get(key)
            if g is None or (g is not False and g["sig"] != self._signature()):
                g = self._graphs[key] = self._capture(key, pts, gts, batch, lods)
        if g:
            g["pts"].copy_(pts.reshape(g["pts"].shape))
            g["gts"].copy_(gts.reshape(g["gts"].shape))
            g["graph"].replay()
        else:
            self._compute_grads(pts, gts, batch, lods)
        self._reduce_and_update()
        net.mark_grids_dirty()      # the Adam kernel writes the parameters behind torch's version counters
        return self.loss


# ------------------------------------------------------------------------------------------------ Trainer
import logging as log
import os
from datetime import datetime


class Trainer(object):
    """The reference's training loop (sdf-net/lib/trainer.py:89-502) around the fused step: same constructor
    `(args, args_str)`, same hooks (`set_dataset/set_network/set_optimizer/set_renderer/set_logger`, `pre_epoch`,
    `grow`, `iterate`, `step_geometry`, `post_epoch`, `log_tb`, `save_model`, `resample`, `train`), same LOD growth
    strategies, resample cadence, loss bookkeeping and checkpoint naming -- so `app/main.py` runs unchanged.

    What differs is where things live: the dataset stays on the device (no DataLoader / pinned-memory hop per
    512-point batch), a step is the fused forward+loss+backward kernels + one Adam launch over flat buffers, losses
    are accumulated on the device and read once per epoch (the reference calls `.item()` twice per iteration), and
    with several ranks each takes a slice of every batch and the flat gradient is exchanged once per step (reduce-scatter +
    Adam on 1/N of the parameters + all-gather; `FusedTrainer(sharded=False)`: one all-reduce).
    TensorBoard is used when importable; its absence only disables the image/scalar summaries.
    """

    def __init__(self, args, args_str=""):
        self.args = args
        self.args_str = args_str
        self.args.epochs += 1                                     # trainer.py:99
        if not torch.cuda.is_available():
            raise RuntimeError("training needs a CUDA device (nglod_b200 has no CPU path)")
        self.rank, self.world, local = ndist.init_from_env()
        self.device = torch.device("cuda", local)
        torch.cuda.set_device(self.device)
        self.latents = None
        self.dataset_size = None
        self.log_dict = {}
        self.loss_lods = list(range(self.args.num_lods))
        self.set_dataset()
        self.set_network()
        self.set_optimizer()
        self.set_renderer()
        self.set_logger()

    # -- construction hooks
    def set_dataset(self):
        from .datasets import MeshDataset
        classes = {"MeshDataset": MeshDataset}
        self.train_dataset = classes[self.args.mesh_dataset](self.args, device=self.device)
        log.info("Dataset Size: {}".format(len(self.train_dataset)))

    def set_network(self):
        from . import models
        self.net = getattr(models, self.args.net)(self.args)
        if self.args.pretrained:
            self.net.load_state_dict(torch.load(self.args.pretrained, map_location="cpu"))
        self.net.to(self.device)
        if self.world > 1:                                        # replicas start identical
            for p in self.net.parameters():
                torch.distributed.broadcast(p.data, src=0)
        log.info("Total number of parameters: {}".format(sum(p.numel() for p in self.net.parameters())))

    def set_optimizer(self):
        if self.args.optimizer == "adam" and getattr(self.net, "padded", False):
            # a model smaller than the kernels' 32 / 128 runs zero-padded: autograd step (same kernels), torch's Adam
            self.optimizer = torch.optim.Adam(self.net.parameters(), lr=self.args.lr)
        elif self.args.optimizer == "adam":
            self.optimizer = FusedTrainer(self.net, lr=self.args.lr)
        elif self.args.optimizer == "sgd":
            self.optimizer = torch.optim.SGD(self.net.parameters(), lr=self.args.lr, momentum=0.8)
        else:
            raise ValueError("Invalid optimizer.")

    def set_renderer(self):
        from . import tracer as tracers
        from .renderer import Renderer
        self.log_tracer = getattr(tracers, self.args.tracer)(self.args)
        self.renderer = Renderer(self.log_tracer, args=self.args, device=self.device)

    def set_logger(self):
        self.log_fname = self.args.exp_name if self.args.exp_name else datetime.now().strftime("%Y%m%d-%H%M%S")
        self.log_dir = os.path.join(self.args.logs, self.log_fname)
        self.writer = None
        if self.rank == 0:
            try:
                from torch.utils.tensorboard import SummaryWriter
                self.writer = SummaryWriter(self.log_dir, purge_step=0)
                self.writer.add_text("Parameters", self.args_str)
            except Exception:  # noqa: BLE001  -- tensorboard is optional
                self.writer = None

    # -- epoch hooks
    def pre_epoch(self, epoch):
        self.loss_lods = list(range(0, self.args.num_lods))
        if self.args.grow_every > 0:
            self.grow(epoch)
        if self.args.only_last:
            self.loss_lods = self.loss_lods[-1:]
        if epoch % self.args.resample_every == 0:
            self.resample(epoch)
        if epoch == self.args.freeze:
            self.net.freeze()
        self.net.train()
        self._epoch_loss = torch.zeros(2, device=self.device)      # [last-LOD l2 sum, total sum] (un-normalised)
        self.log_dict["l2_loss"] = 0
        self.log_dict["total_loss"] = 0
        self.log_dict["total_iter_count"] = 0

    def grow(self, epoch):
        """Coarse-to-fine schedules (trainer.py:262-276)."""
        n = self.args.num_lods
        stage = min(n, (epoch // self.args.grow_every) + 1)
        every = list(range(0, n))
        strategy = self.args.growth_strategy
        if strategy == "onebyone":
            self.loss_lods = [stage - 1]
        elif strategy == "increase":
            self.loss_lods = every[:stage]
        elif strategy == "shrink":
            self.loss_lods = every[stage - 1:]
        elif strategy == "finetocoarse":
            self.loss_lods = every[n - stage:]
        elif strategy == "onlylast":
            self.loss_lods = every[-1:]
        else:
            raise NotImplementedError

    def batches(self):
        """Shuffled batches of the device-resident dataset (what DataLoader(shuffle=True) yields in the reference);
        with several ranks every rank walks the same permutation and takes its slice of each batch."""
        n = len(self.train_dataset)
        g = torch.Generator(device=self.device)
        g.manual_seed(int(torch.randint(0, 2 ** 31 - 1, (1,)).item()) if self.world == 1 else 1234 + self._epoch)
        perm = torch.randperm(n, device=self.device, generator=g)
        bs = self.args.batch_size
        for start in range(0, n, bs):
            idx = perm[start:start + bs]
            s, e = ndist.shard_range(idx.shape[0], self.rank, self.world)
            yield self.train_dataset.pts[idx[s:e]], self.train_dataset.d[idx[s:e]], idx.shape[0]

    def iterate(self, epoch):
        self._epoch = epoch
        self.dataset_size = (len(self.train_dataset) + self.args.batch_size - 1) // self.args.batch_size
        for n_iter, data in enumerate(self.batches()):
            self.step_geometry(epoch, n_iter, data)

    def step_geometry(self, epoch, n_iter, data):
        """loss = sum_{lod in loss_lods} sum_i (sdf_lod(x_i) - gt_i)^2 / batch; backward; optimizer step
        (trainer.py:294-340).  `data` = (pts, gts, global batch size)."""
        pts, gts, batch = data
        if isinstance(self.optimizer, FusedTrainer):
            self.optimizer.step(pts, gts, global_batch=batch, loss_lods=self.loss_lods)
            lod_loss = self.optimizer.lod_loss
            self._epoch_loss[0] += lod_loss[self.loss_lods[-1]] * batch
            self._epoch_loss[1] += self.optimizer.loss[0] * batch
        else:
            self.net.zero_grad()
            loss, last = 0, None
            for lod in self.loss_lods:
                last = ((self.net.sdf(pts, lod=lod) - gts) ** 2).sum()
                loss = loss + last
            self._epoch_loss[0] += last.detach()
            self._epoch_loss[1] += loss.detach()
            (loss / batch).backward()
            if self.world > 1:
                for p in self.net.parameters():
                    if p.grad is not None:
                        torch.distributed.all_reduce(p.grad)
            self.optimizer.step()
        self.log_dict["total_iter_count"] += batch

    def post_epoch(self, epoch):
        self.net.eval()
        self.log_tb(epoch)
        if epoch % self.args.save_every == 0 and self.rank == 0:
            self.save_model(epoch)
        if epoch % self.args.render_every == 0 and self.writer is not None:
            self.render_tb(epoch)

    def log_tb(self, epoch):
        sums = self._epoch_loss.clone()
        ndist.allreduce_sum_(sums)
        l2, total = (float(v) for v in sums.cpu())               # the one host read of the epoch
        self.log_dict["l2_loss"] = l2 / (self.log_dict["total_iter_count"] + 1e-6)
        self.log_dict["total_loss"] = total / (self.log_dict["total_iter_count"] + 1e-6)
        log.info("EPOCH {}/{} | total loss: {:>.3E} | l2 loss: {:>.3E}".format(
            epoch + 1, self.args.epochs, self.log_dict["total_loss"], self.log_dict["l2_loss"]))
        if self.writer is not None:
            self.writer.add_scalar("Loss/l2_loss", self.log_dict["l2_loss"], epoch)
            self.writer.add_scalar("Loss/total_loss", self.log_dict["total_loss"], epoch)

    def render_tb(self, epoch):
        for d in range(self.args.num_lods):
            self.net.lod = d
            out = self.renderer.shade_images(self.net, f=self.args.camera_origin, t=self.args.camera_lookat,
                                             fov=self.args.camera_fov).image().byte().numpy()
            for tag, img in (("Depth", out.depth), ("Hit", out.hit), ("Normal", out.normal), ("RGB", out.rgb)):
                self.writer.add_image(f"{tag}/{d}", img.transpose(2, 0, 1), epoch)
            self.net.lod = None

    def save_model(self, epoch):
        """Checkpoint naming of trainer.py:416-442: <model_path>/<exp>[-<epoch>].pth, state_dict unless --save-all."""
        parts = self.log_fname.split("/")
        os.makedirs(os.path.join(self.args.model_path, *parts[:-1]), exist_ok=True)
        name = f"{self.log_fname}-{epoch}.pth" if self.args.save_as_new else f"{self.log_fname}.pth"
        fname = os.path.join(self.args.model_path, name)
        log.info(f"Saving model checkpoint to: {fname}")
        if self.args.save_all:
            torch.save(self.net, fname)
        else:   # contiguous NCDHW copies: the file is byte-compatible with a reference checkpoint
            torch.save({k: v.detach().cpu().contiguous() for k, v in self.net.state_dict().items()}, fname)
        return fname

    def resample(self, epoch):
        self.train_dataset.resample()

    def train(self):
        for epoch in range(self.args.epochs):
            self.pre_epoch(epoch)
            self.iterate(epoch)
            self.post_epoch(epoch)
        if self.writer is not None:
            self.writer.close()
