"""CPU: pin the oracle (oracle/nglod_oracle.py) against vectors produced by the unmodified reference
(tests/golden/make_golden.py).  If these pass, the oracle may be trusted as the checker for the CUDA path."""
import numpy as np
import torch

from oracle import nglod_oracle as O
from helpers import rand5_model, fit3_model, weights_checksum, cl_flat


def test_same_seed_same_weights(rand5):
    net, _ = rand5_model()
    np.testing.assert_allclose(weights_checksum(net), rand5["weights_checksum"], rtol=1e-12)


def test_state_dict_contract():
    net, _ = rand5_model()
    sd = net.state_dict()
    for i, r in enumerate([4, 8, 16, 32, 64]):
        assert tuple(sd[f"features.{i}.fm"].shape) == (1, 32, r + 1, r + 1, r + 1)
        assert tuple(sd[f"louts.{i}.0.weight"].shape) == (128, 35)
        assert tuple(sd[f"louts.{i}.2.weight"].shape) == (1, 128)
    assert sum(p.numel() for p in net.parameters()) == 10146213


def test_oracle_sdf_matches_reference(rand5):
    net, _ = rand5_model()
    onet = O.OracleNet(net.state_dict())
    x = torch.from_numpy(rand5["x"])
    for l in range(5):
        d = onet.sdf(x, lod=l).numpy()
        assert np.abs(d - rand5[f"sdf_lod{l}"]).max() < 1e-6
    lst = onet.sdf(x, return_lst=True)
    assert np.abs(np.stack([t.numpy() for t in lst]) - rand5["sdf_lst"]).max() < 1e-6
    onet.lod = 3
    assert np.abs(onet(x).numpy() - rand5["forward_lod3"]).max() < 1e-6


def test_explicit_f64_restatement_matches_reference(rand5, fit3):
    net, _ = rand5_model()
    onet = O.OracleNet(net.state_dict())
    x = torch.from_numpy(rand5["x"])
    for l in (0, 2, 4):
        assert np.abs(O.sdf_explicit_f64(onet, x, l) - rand5[f"sdf_lod{l}"]).max() < 2e-6
    net3, _ = fit3_model(fit3)
    o3 = O.OracleNet(net3.state_dict())
    x3 = torch.from_numpy(fit3["x"])
    for l in range(3):
        assert np.abs(O.sdf_explicit_f64(o3, x3, l) - fit3[f"sdf_lod{l}"]).max() < 2e-6
        assert np.abs(o3.sdf(x3, lod=l).numpy() - fit3[f"sdf_lod{l}"]).max() < 1e-6


def test_oracle_gradients_match_reference(rand5):
    net, _ = rand5_model()
    onet = O.OracleNet(net.state_dict(), requires_grad=True)
    x = torch.from_numpy(rand5["x"])
    gt = torch.from_numpy(rand5["gt"])
    for tag, lods in (("g4", [4]), ("g13", [1, 3])):
        loss = O.l2_loss_and_grads(onet, x, gt, lods)
        assert abs(loss.item() - float(rand5[f"{tag}_loss"])) < 1e-6
        for i in range(max(lods) + 1):
            g = cl_flat(onet.fm[i].grad).numpy()
            scale = np.abs(g).max() + 1e-30
            if i <= 1:
                assert np.abs(g - rand5[f"{tag}_fm{i}"]).max() / scale < 1e-5
            else:
                assert np.abs(g[rand5[f"{tag}_fm{i}_idx"]] - rand5[f"{tag}_fm{i}_val"]).max() / scale < 1e-5
        for l in lods:
            for k, t in zip(("0.weight", "0.bias", "2.weight", "2.bias"), onet.dec[l]):
                ref = rand5[f"{tag}_louts{l}.{k}"]
                assert np.abs(t.grad.numpy() - ref).max() / (np.abs(ref).max() + 1e-30) < 1e-5
        assert onet.dec[0][0].grad is None or float(onet.dec[0][0].grad.abs().sum()) == 0.0


def _check_trace(res, gold, prefix):
    hit = gold[prefix + "_hit"]
    assert np.array_equal(res["hit"].numpy(), hit)
    assert np.abs(res["depth"].numpy() - gold[prefix + "_depth"])[hit].max(initial=0) < 1e-5
    assert np.abs(res["x"].numpy() - gold[prefix + "_x"])[hit].max(initial=0) < 1e-5
    assert np.abs(res["normal"].numpy() - gold[prefix + "_normal"]).max() < 1e-4


def test_oracle_tracer_matches_reference(rand5, fit3):
    net, _ = rand5_model()
    onet = O.OracleNet(net.state_dict())
    onet.lod = 4
    res = O.sphere_trace(onet, torch.from_numpy(rand5["t1_ray_o"]), torch.from_numpy(rand5["t1_ray_d"]))
    _check_trace(res, rand5, "t1")
    res = O.sphere_trace(onet, torch.from_numpy(rand5["t2_ray_o"]), torch.from_numpy(rand5["t2_ray_d"]), num_steps=12)
    _check_trace(res, rand5, "t2")
    net3, _ = fit3_model(fit3)
    o3 = O.OracleNet(net3.state_dict())
    o3.lod = 2
    res = O.sphere_trace(o3, torch.from_numpy(fit3["t1_ray_o"]), torch.from_numpy(fit3["t1_ray_d"]))
    _check_trace(res, fit3, "t1")


def test_oracle_look_at_matches_reference(rand5):
    torch.manual_seed(123)
    o, d = O.look_at([-2.8, 2.8, -2.8], [0, 0, 0], 64, 36, fov=30.0)
    assert np.array_equal(o.numpy(), rand5["t1_ray_o"])
    assert np.abs(d.numpy() - rand5["t1_ray_d"]).max() < 1e-7


def test_oracle_finitediff_matches_reference(rand5):
    net, _ = rand5_model()
    onet = O.OracleNet(net.state_dict())
    onet.lod = 4
    x = torch.from_numpy(rand5["x"][:512])
    with torch.no_grad():
        g = O.gradient_finitediff(x, onet).numpy()
    assert np.abs(g - rand5["finitediff_lod4"]).max() < 1e-4      # differences of ~1e-7 values divided by 2h


def test_oracle_aabb_semantics():
    o = torch.tensor([[-2.8, 2.8, -2.8], [0.1, 0.2, 0.3], [3.0, 0.0, 0.0], [3.0, 0.0, 0.0], [0.0, 5.0, 0.0]])
    d = torch.tensor([[1.0, -1.0, 1.0], [1.0, 0.0, 0.0], [-1.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])
    d = torch.nn.functional.normalize(d, dim=1)
    x, t, hit = O.aabb(o, d)
    assert hit.tolist() == [True, False, True, False, True]
    assert torch.equal(x[1], o[1]) and t[1].item() == 0.0           # origin inside: untouched defaults
    assert abs(t[2].item() - 2.0) < 1e-6 and abs(x[2, 0].item() - 1.0) < 1e-6
    assert torch.equal(x[3], o[3]) and t[3].item() == 0.0           # pointing away


def test_oracle_mesh2sdf_on_sphere():
    from nglod_b200.lib.torchgp import icosphere
    V, F = icosphere(2)
    g = torch.Generator().manual_seed(0)
    p = torch.rand(400, 3, generator=g) * 2 - 1
    d = O.mesh2sdf(p, V[F])
    r = p.norm(dim=1)
    far = (r - 1).abs() > 0.06                                      # outside the faceting band
    assert torch.equal(d[far] < 0, (r < 1)[far])
    assert ((d - (r - 1)).abs() < 0.06).all()


def test_philox_known_answers():
    """Random123's published known-answer vectors for philox4x32-10 (kat_vectors): the generator the sampler kernel uses."""
    assert O.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert O.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert O.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
