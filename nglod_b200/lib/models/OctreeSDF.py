"""OctreeSDF -- the reference's public model class (sdf-net/lib/models/OctreeSDF.py:59-155)
on top of the fused sm_100a kernels.

Same constructor arguments, parameter names, shapes and init order as the
reference (`features.{i}.fm` [1,F,R+1,R+1,R+1] with R = 2^(i+base_lod) drawn as
randn*0.01, then `louts.{i}` = Linear(3+F,H)-ReLU-Linear(H,1)), so a reference
checkpoint loads unchanged and the same seed gives the same weights.  What is
different is how it runs: the grids are kept in torch.channels_last_3d memory
format (one corner = one 128-byte line) and `sdf()` is ONE fused kernel
(multi-LOD trilinear gather + running sum + decoder), with a recompute-based
backward -- instead of the per-LOD grid_sample / cat / Linear launch chain.
"""
import weakref

import torch
import torch.nn as nn
from torch.autograd.function import once_differentiable

from .BaseLOD import BaseLOD
from ... import ops, _lib


KERNEL_F, KERNEL_H = 32, 128     # the shape libnglod_b200.so is built for (csrc/common.cuh NGLOD_F / NGLOD_H)


def _pad_grid(fm):
    """[1,F,S,S,S] channels-last -> [1,32,S,S,S] channels-last, channels F.. zero."""
    out = torch.zeros(1, KERNEL_F, *fm.shape[2:], device=fm.device, dtype=fm.dtype).contiguous(memory_format=torch.channels_last_3d)
    out[:, :fm.shape[1]] = fm
    return out


class FeatureVolume(nn.Module):
    """Dense (fsize+1)^3 x fdim feature grid (reference: OctreeSDF.py:38-57)."""

    def __init__(self, fdim, fsize):
        super().__init__()
        self.fsize = fsize
        self.fdim = fdim
        fm = torch.randn(1, fdim, fsize + 1, fsize + 1, fsize + 1) * 0.01
        self.fm = nn.Parameter(fm.contiguous(memory_format=torch.channels_last_3d))
        self.sparse = None

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        if not self.fm.data.is_contiguous(memory_format=torch.channels_last_3d):
            self.fm.data = self.fm.data.contiguous(memory_format=torch.channels_last_3d)
        return out

    def forward(self, x):
        """Trilinear sample of this LOD alone: [N,3] -> [N,fdim] ([N,K,3] -> [N,K,fdim])."""
        if self.fdim > KERNEL_F:
            raise RuntimeError(f"feature-dim {self.fdim}: the kernels are built for at most {KERNEL_F} channels")
        view = ops.NetView.grids_only([self.fm.data if self.fdim == KERNEL_F else _pad_grid(self.fm.data)])
        return ops.sdf_features(view, 0, x.reshape(-1, 3))[:, :self.fdim].reshape(*x.shape[:-1], self.fdim)


class _SdfFunction(torch.autograd.Function):
    """d = sdf(x, lod); backward recomputes the forward in-kernel (no saved activations)."""

    @staticmethod
    def forward(ctx, x, module, lod, *params):
        view = module.net_view(inference=False)
        ctx.module, ctx.lod = module, lod
        ctx.save_for_backward(x)
        return ops.sdf_forward(view, lod, x)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        (x,) = ctx.saved_tensors
        module, lod = ctx.module, ctx.lod
        view = module.net_view(inference=False)
        needs = ctx.needs_input_grad
        n_grids = lod + 1
        # gradients in the kernels' shapes (== the parameters' shapes unless the model is zero-padded, see OctreeSDF.padded)
        grid_grads = [torch.zeros_like(view.grids[i], memory_format=torch.preserve_format)
                      if needs[3 + i] else None for i in range(n_grids)]
        dec_grads = [torch.zeros_like(p) if needs[3 + n_grids + k] else None for k, p in enumerate(view.decoders[lod])]
        scratch = None
        if view.summed is not None:       # fresh zero gradients double as the scratch (aliased: no copy-out pass)
            own = module.summed_grad_scratch()
            scratch = [g if g is not None else own[i] for i, g in enumerate(grid_grads)] + own[n_grids:]
        gx = ops.sdf_backward(view, lod, x, grad_out.contiguous(), grid_grads + [None] * (view.num_lods - n_grids),
                              tuple(dec_grads), want_grad_x=needs[0], summed_scratch=scratch,
                              scatter_scratch=module.scatter_scratch() if view.summed is not None else None)
        if module.padded:
            grid_grads, dec_grads = module._unpad_grads(grid_grads, dec_grads)
        return (gx, None, None, *grid_grads, *dec_grads)


# net -> {(inference, use_summed): (key, NetView)}: the borrowed views of net_view(), see there
_VIEWS = weakref.WeakKeyDictionary()


class OctreeSDF(BaseLOD):
    def __init__(self, args, init=None):
        super().__init__(args)
        self.fdim = self.args.feature_dim
        self.fsize = self.args.feature_size
        self.hidden_dim = self.args.hidden_dim
        self.pos_invariant = self.args.pos_invariant
        # The library is built for feature-dim 32 / hidden-dim 128 (the reference's defaults and every published
        # configuration).  SMALLER models (`--feature-dim 16`, README.md:109 of the reference) run through the same kernels
        # zero-padded: channels F..31 of the grids and the matching W0 columns are zero, hidden units H..127 have
        # W0 = b0 = W1 = 0 -- the padded net is the same function exactly (every extra term is 0 * 0 or W1 = 0 times
        # relu(0)), and its gradients restricted to the real entries are the real gradients.  Larger shapes raise.
        if self.fdim > KERNEL_F or self.hidden_dim > KERNEL_H:
            raise RuntimeError(f"feature-dim {self.fdim} / hidden-dim {self.hidden_dim}: the kernels are built for at most "
                               f"{KERNEL_F} / {KERNEL_H} (smaller models are zero-padded, larger ones are not supported)")
        self.padded = self.fdim != KERNEL_F or self.hidden_dim != KERNEL_H
        self._padded_cache = None

        self.features = nn.ModuleList(
            [FeatureVolume(self.fdim, 2 ** (i + self.args.base_lod)) for i in range(self.args.num_lods)])
        self.interpolate = self.args.interpolate
        # how the 35->128 contraction runs: "tc" = tcgen05 3xTF32 (|err| ~1e-6), "fp32" = CUDA cores (|err| ~1e-7)
        self.math_mode = getattr(args, "math_mode", None) or "tc"
        # How the inference kernels (no-grad sdf(), finite-difference normals, the tracers) read the grids.
        #   sum_lods     True: they gather ONE grid -- the prefix sum of LODs 0..lod resampled at LOD lod's nodes.  The
        #                LOD grids nest, so this is the same function as the reference's running sum of per-LOD samples
        #                (OctreeSDF.py:109-110) exactly in real arithmetic, to fp32 rounding (~1e-8) in practice.
        #                False: gather every LOD (what training / autograd always does).
        #   grid_storage "fp32": summed grids in fp32;  "fp16": also as packed half-precision "x-pair lines" for the
        #                tensor-core kernels (half the gather traffic; equals the fp32 kernels run on the fp16-rounded
        #                summed grid; the reference's real-time export stores features as fp16 too, SOL_NGLOD.py:73).
        # The derived copies are rebuilt whenever a grid changes (torch version counters; mark_grids_dirty() after
        # writes that bypass them).
        self.sum_lods = getattr(args, "sum_lods", None)
        self.sum_lods = True if self.sum_lods is None else bool(self.sum_lods)
        self.grid_storage = getattr(args, "grid_storage", None) or "fp32"
        self._derived = None            # (key, summed list, half list)

        self.sdf_input_dim = self.fdim + (0 if self.pos_invariant else self.input_dim)
        self.num_decoder = 1 if args.joint_decoder else self.args.num_lods
        self.louts = nn.ModuleList([
            nn.Sequential(nn.Linear(self.sdf_input_dim, self.hidden_dim, bias=True), nn.ReLU(),
                          nn.Linear(self.hidden_dim, 1, bias=True))
            for _ in range(self.num_decoder)])

    # ------------------------------------------------------------------ kernel plumbing
    def decoder_params(self, lod):
        seq = self.louts[0 if self.num_decoder == 1 else lod]
        return (seq[0].weight, seq[0].bias, seq[2].weight, seq[2].bias)

    # ------------------------------------------------------------------ zero-padding to the kernels' shape
    def _kernel_params(self):
        """(grids, decoders) as the kernels take them: the parameters themselves, or -- for a model smaller than 32 / 128 --
        zero-padded copies, rebuilt when a parameter was written (version counters; mark_grids_dirty())."""
        grids = [f.fm.data for f in self.features]
        decs = [tuple(p.data for p in self.decoder_params(i)) for i in range(self.num_lods)]
        if not self.padded:
            return grids, decs
        key = [(p._version, p.data_ptr()) for p in self.parameters()]
        if self._padded_cache is None or self._padded_cache[0] != key:
            F_, H_ = self.fdim, self.hidden_dim
            xyz = 0 if self.pos_invariant else 3
            pg = [_pad_grid(g) for g in grids]
            pd = []
            for i in range(self.num_decoder):
                w0, b0, w1, b1 = (p.data for p in self.decoder_params(i))
                w0p = torch.zeros(KERNEL_H, xyz + KERNEL_F, device=w0.device)
                w0p[:H_, :xyz + F_] = w0
                b0p = torch.zeros(KERNEL_H, device=w0.device); b0p[:H_] = b0
                w1p = torch.zeros(1, KERNEL_H, device=w0.device); w1p[:, :H_] = w1
                pd.append((w0p, b0p, w1p, b1.clone()))
            self._padded_cache = [key, pg, pd]
        pd = self._padded_cache[2]
        return self._padded_cache[1], [pd[0 if self.num_decoder == 1 else i] for i in range(self.num_lods)]

    def _unpad_grads(self, grid_grads, dec_grads):
        F_, H_ = self.fdim, self.hidden_dim
        xyz = 0 if self.pos_invariant else 3
        gg = [None if g is None else g[:, :F_].contiguous(memory_format=torch.channels_last_3d) for g in grid_grads]
        gw0, gb0, gw1, gb1 = dec_grads
        dg = [None if gw0 is None else gw0[:H_, :xyz + F_].contiguous(), None if gb0 is None else gb0[:H_].contiguous(),
              None if gw1 is None else gw1[:, :H_].contiguous(), gb1]
        return gg, dg

    def _grids_nest(self):
        res = [f.fsize for f in self.features]
        return all(res[i] % res[j] == 0 for i in range(len(res)) for j in range(i))

    def mark_grids_dirty(self):
        """MUST be called after writing the grids behind torch's back -- `fm.data.copy_(...)`, raw pointers, a custom
        kernel such as the fused Adam step: such writes do not bump the tensors' version counters, and the inference
        kernels read DERIVED prefix-summed grids that are rebuilt only when a version counter or an address changes.
        `load_state_dict` and device / dtype moves call it themselves.  The derived buffers are kept (a captured CUDA
        graph of the training step rebuilds them in place)."""
        if self._derived is not None:
            self._derived[0] = None
        self._padded_cache = None
        _VIEWS.pop(self, None)

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.mark_grids_dirty()
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        if getattr(self, "_derived", None) is not None:
            self._derived = None            # moved / cast: the derived buffers belong to the old placement
        _VIEWS.pop(self, None)
        return out

    def _derived_grids(self, want_half=False):
        """(summed, summed_half) rebuilt when a grid was written or moved; the fp16 copy only when asked for."""
        key = [(f.fm._version, f.fm.data_ptr()) for f in self.features]
        if self._derived is None or self._derived[0] != key:
            grids = self._kernel_params()[0]
            base = ops.NetView.grids_only(grids)
            old = self._derived[1] if self._derived is not None else [None] * len(grids)
            old_half = self._derived[2] if self._derived is not None else None
            summed = []
            for i, (g, o) in enumerate(zip(grids, old)):      # ascending: level i = prolongation of level i-1 + grid i
                if i > 0:
                    base.struct.summed[i - 1] = summed[i - 1].data_ptr()
                summed.append(ops.build_summed_grid(base, i, out=o if (o is not None and o.device == g.device) else None))
            self._derived = [key, summed, None, old_half]
        if want_half and self._derived[2] is None:
            oh = self._derived[3] if len(self._derived) > 3 and self._derived[3] is not None else [None] * len(self._derived[1])
            self._derived[2] = [ops.pack_grid_fp16(sg, out=o if (o is not None and o.device == sg.device) else None)
                                for sg, o in zip(self._derived[1], oh)]
            self._derived[3] = None
        return self._derived[1], (self._derived[2] if want_half else None)

    def summed_grad_scratch(self):
        """Zero-filled buffers shaped like the grids for the single-grid backward (nglod_net_grad_t.summed); the
        kernels leave them zero, so they are allocated once."""
        sc = getattr(self, "_summed_scratch", None)
        if sc is None or any(t.device != f.fm.device for t, f in zip(sc, self.features)):
            sc = [torch.zeros_like(g, memory_format=torch.preserve_format) for g in self._kernel_params()[0]]
            self._summed_scratch = sc
        return sc

    def scatter_scratch(self):
        """Zero-filled 1 MB buffer for the private scatter copies of the small grids (nglod_net_grad_t.scatter_scratch);
        the kernels leave it zero, so it is allocated once."""
        sc = getattr(self, "_scatter_scratch", None)
        dev = self.features[0].fm.device
        if sc is None or sc.device != dev:
            sc = torch.zeros(1 << 18, dtype=torch.float32, device=dev)
            self._scatter_scratch = sc
        return sc

    def summed_state_dict(self, lod=None):
        """state_dict of the function the inference kernels evaluate for sdf(x, lod), in the reference's own format:
        every `features.i.fm` zero except `features.{lod}.fm`, which holds the prefix-summed grid (rounded through fp16
        when grid_storage == 'fp16').  A reference OctreeSDF loaded with it returns this model's sdf(x, lod)."""
        lod = self.num_lods - 1 if lod is None else lod
        summed, _ = self._derived_grids()
        sd = {k: v.detach().clone() for k, v in self.state_dict().items()}
        for i in range(self.num_lods):
            sd[f"features.{i}.fm"] = torch.zeros_like(sd[f"features.{i}.fm"])
        top = summed[lod][:, :self.fdim].detach().clone()
        sd[f"features.{lod}.fm"] = top.half().float() if self.grid_storage == "fp16" else top
        return sd

    def net_view(self, inference=True, use_summed=True):
        """Borrow the current parameters as an nglod_net_t.  inference=False (the autograd / training kernels) never
        attaches the half-precision copy; use_summed=False leaves the prefix-summed grids out (small training batches:
        rebuilding them costs more than five short gathers).  The view is re-used while no parameter was written or
        moved (version counters + addresses, the same key the derived grids use; mark_grids_dirty() drops it): building
        one costs ~50 us of host time, which a 1 ms frame notices."""
        ps = [f._parameters["fm"] for f in self.features._modules.values()]     # direct: Module.parameters() costs 25 us
        for seq in self.louts._modules.values():
            for lin in (seq._modules["0"], seq._modules["2"]):
                ps.append(lin._parameters["weight"])
                ps.append(lin._parameters["bias"])
        key = (self.math_mode, self.grid_storage, self.sum_lods, self.pos_invariant,
               tuple([(p._version, p.data_ptr()) for p in ps]))
        views = _VIEWS.setdefault(self, {})     # kept off the module: a view holds ctypes pointers (no deepcopy / pickle)
        hit = views.get((inference, use_summed))
        if hit is not None and hit[0] == key:
            return hit[1]
        grids, decs = self._kernel_params()
        summed = half = None
        if use_summed and self.sum_lods and self._grids_nest():
            summed, half = self._derived_grids(want_half=inference and self.grid_storage == "fp16" and self.math_mode == "tc")
        view = ops.NetView(grids, decs, pos_invariant=self.pos_invariant,
                           math_mode=_lib.MATH_TC3XTF32 if self.math_mode == "tc" else _lib.MATH_FP32,
                           summed=summed, summed_half=half)
        views[(inference, use_summed)] = (key, view)
        return view

    def _eval_lod(self, x, lod):
        shape = x.shape
        x2 = x.reshape(-1, 3)
        if x2.dtype != torch.float32:
            x2 = x2.float()
        if torch.is_grad_enabled() and (x2.requires_grad or any(p.requires_grad for p in self.parameters())):
            params = [self.features[i].fm for i in range(lod + 1)] + list(self.decoder_params(lod))
            d = _SdfFunction.apply(x2, self, lod, *params)
        else:
            d = ops.sdf_forward(self.net_view(), lod, x2)
        return d.reshape(*shape[:-1], 1)

    # ------------------------------------------------------------------ reference API
    def encode(self, x):
        return x   # encoding is disabled for OctreeSDF (OctreeSDF.py:90-92)

    def sdf(self, x, lod=None, return_lst=False):
        """Reference semantics (OctreeSDF.py:94-155): with an integer lod return that head's
        distance [N,1]; with lod None (and self.lod None) evaluate every head, stash them in
        `loss_preds` while training, and return the list or its last element."""
        if lod is None:
            lod = self.lod
        if self.interpolate is not None and lod is not None:
            raise NotImplementedError("--interpolate is broken in the reference (NameError at OctreeSDF.py:138)")
        if lod is not None and 0 <= lod < self.num_lods:
            return self._eval_lod(x, lod)
        preds = [self._eval_lod(x, i) for i in range(self.num_lods)]
        if self.training:
            self.loss_preds = preds
        return preds if return_lst else preds[-1]
