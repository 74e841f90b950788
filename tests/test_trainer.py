"""Trainer loop and the .npz export (8f-2)."""
import os
import tempfile
import types

import numpy as np
import pytest
import torch

from helpers import make_args


def test_growth_strategies_match_reference_table():
    """trainer.py:262-276 for num_lods = 4, grow_every = 2."""
    from nglod_b200.lib.trainer import Trainer
    t = Trainer.__new__(Trainer)
    expect = {
        "onebyone": {0: [0], 2: [1], 5: [2], 6: [3], 9: [3]},
        "increase": {0: [0], 2: [0, 1], 4: [0, 1, 2], 6: [0, 1, 2, 3], 9: [0, 1, 2, 3]},
        "shrink": {0: [0, 1, 2, 3], 2: [1, 2, 3], 4: [2, 3], 6: [3]},
        "finetocoarse": {0: [3], 2: [2, 3], 4: [1, 2, 3], 6: [0, 1, 2, 3]},
        "onlylast": {0: [3], 6: [3]},
    }
    for strategy, table in expect.items():
        t.args = types.SimpleNamespace(num_lods=4, grow_every=2, growth_strategy=strategy)
        for epoch, lods in table.items():
            t.grow(epoch)
            assert t.loss_lods == lods, (strategy, epoch)
    t.args.growth_strategy = "nope"
    with pytest.raises(NotImplementedError):
        t.grow(0)


@pytest.mark.gpu
def test_train_checkpoint_export_roundtrip():
    from nglod_b200.lib.trainer import Trainer
    from nglod_b200.lib.models import OctreeSDF
    from nglod_b200.lib.torchgp import icosphere, write_obj
    from nglod_b200.lib import spc as S
    with tempfile.TemporaryDirectory() as td:
        V, F = icosphere(3)
        obj = os.path.join(td, "ball.obj")
        write_obj(obj, V * 0.7, F)
        args = make_args(["--num-lods", "3", "--dataset-path", obj, "--epochs", "7", "--batch-size", "2048",
                          "--num-samples", "8000", "--exp-name", "unit/ball", "--model-path", os.path.join(td, "models"),
                          "--logs", os.path.join(td, "logs"), "--render-every", "100", "--resample-every", "2",
                          "--grow-every", "1", "--growth-strategy", "increase", "--render-res", "64", "64"])
        torch.manual_seed(0)
        tr = Trainer(args, "```args```")
        assert args.epochs == 8                                   # the reference's off-by-one (trainer.py:99)
        losses = []
        for epoch in range(args.epochs):
            tr.pre_epoch(epoch)
            assert tr.loss_lods == list(range(min(3, epoch + 1)))
            tr.iterate(epoch)
            tr.post_epoch(epoch)
            losses.append(tr.log_dict["l2_loss"])
        assert tr.dataset_size == 20 and tr.log_dict["total_iter_count"] == 40000
        assert losses[-1] < 0.5 * losses[2]                         # same loss_lods from epoch 2 on
        ckpt = os.path.join(td, "models", "unit", "ball.pth")
        sd = torch.load(ckpt)
        assert all(v.is_contiguous() for v in sd.values())         # reference-layout file
        net2 = OctreeSDF(make_args(["--num-lods", "3"]))
        net2.load_state_dict(sd)
        x = torch.rand(1000, 3, device="cuda") * 2 - 1
        with torch.no_grad():
            assert torch.equal(net2.cuda().sdf(x, lod=2), tr.net.sdf(x, lod=2))
            d = tr.net.sdf(x, lod=2)[:, 0]
        assert (d - (x.norm(dim=1) - 1.0)).abs().mean() < 0.1     # the dataset normalises the mesh onto the unit sphere
        # real-time renderer export
        octree = S.mesh_to_octree(tr.train_dataset.V, tr.train_dataset.F, 4, num_samples=1 << 18)
        sp = S.SparseOctreeSDF(tr.net, S.SPC(octree))
        npz = os.path.join(td, "ball.npz")
        sp.save(npz)
        z = np.load(npz)
        assert set(z.files) == {"octree", "cc", "cf", "w0", "b0", "w1", "b1", "pyramid"}
        assert z["octree"].dtype == np.uint8 and z["cc"].dtype == np.uint8 and z["cf"].dtype == np.float16
        assert z["cf"].shape == (int(z["pyramid"].sum()), 32) and z["cc"].shape == (z["cf"].shape[0], 3)
        assert z["w0"].shape == (3, 128, 35) and z["w1"].shape == (3, 1, 128) and z["w0"].dtype == np.float16
        # corner features are the dense grid's values at those corners (LOD 0 rows come first)
        n0 = int(z["pyramid"][0])
        cc0 = torch.from_numpy(z["cc"][:n0].astype(np.int64))
        fm0 = tr.net.features[0].fm.detach().cpu()
        assert np.array_equal(z["cf"][:n0], fm0[0][:, cc0[:, 2], cc0[:, 1], cc0[:, 0]].t().half().numpy())


@pytest.mark.gpu
def test_fused_trainer_respects_requires_grad():
    """`net.freeze()` / requires_grad=False (reference: BaseSDF.freeze, trainer.py:247-251): frozen parameters get no Adam
    update -- not even from stale momentum --, the others keep training; a fully frozen network raises like
    `loss.backward()` does in the reference."""
    from nglod_b200.lib.trainer import FusedTrainer
    from nglod_b200.lib.models import OctreeSDF
    torch.manual_seed(0)
    net = OctreeSDF(make_args(["--num-lods", "3"])).cuda()
    tr = FusedTrainer(net, lr=1e-2)
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand(70001, 3, device="cuda", generator=g) * 2 - 1
    gt = (x.norm(dim=1, keepdim=True) - 0.5)
    for _ in range(3):
        tr.step(x, gt)                                              # builds momentum everywhere
    for f in net.features:
        f.fm.requires_grad_(False)                                  # freeze the grids, keep the decoders
    grids = [f.fm.detach().clone() for f in net.features]
    w = net.louts[2][0].weight.detach().clone()
    for _ in range(2):
        tr.step(x, gt)
    for f, before in zip(net.features, grids):
        assert torch.equal(f.fm.detach(), before)
    assert not torch.equal(net.louts[2][0].weight.detach(), w)
    net.freeze()
    with pytest.raises(RuntimeError):
        tr.step(x, gt)


@pytest.mark.gpu
def test_trainer_smaller_model_trains_through_autograd():
    """`--feature-dim 16 --hidden-dim 64` (the reference README's example of a smaller model): the kernels run it zero-padded,
    the Trainer steps it with torch's Adam through autograd (FusedTrainer needs the kernels' own shape and says so)."""
    from nglod_b200.lib.trainer import Trainer, FusedTrainer
    from nglod_b200.lib.torchgp import icosphere, write_obj
    with tempfile.TemporaryDirectory() as td:
        V, F = icosphere(3)
        obj = os.path.join(td, "ball.obj")
        write_obj(obj, V * 0.7, F)
        args = make_args(["--num-lods", "3", "--feature-dim", "16", "--hidden-dim", "64", "--dataset-path", obj, "--epochs", "5",
                          "--batch-size", "2048", "--num-samples", "8000", "--exp-name", "unit/small",
                          "--model-path", os.path.join(td, "models"), "--logs", os.path.join(td, "logs"),
                          "--render-every", "100", "--lr", "0.005"])
        torch.manual_seed(0)
        tr = Trainer(args, "```args```")
        assert tr.net.padded and isinstance(tr.optimizer, torch.optim.Adam)
        assert tr.net.features[0].fm.shape[1] == 16 and tr.net.louts[0][0].weight.shape == (64, 19)
        with pytest.raises(RuntimeError):
            FusedTrainer(tr.net)
        losses = []
        for epoch in range(args.epochs):
            tr.pre_epoch(epoch)
            tr.iterate(epoch)
            tr.post_epoch(epoch)
            losses.append(tr.log_dict["l2_loss"])
        assert losses[-1] < 0.5 * losses[0]
        x = torch.rand(1000, 3, device="cuda") * 2 - 1
        with torch.no_grad():
            d = tr.net.sdf(x, lod=2)[:, 0]
        assert (d - (x.norm(dim=1) - 1.0)).abs().mean() < 0.15
