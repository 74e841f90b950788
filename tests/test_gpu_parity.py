"""GPU parity tests: the sm_100a kernels (through the C ABI) against
  (1) the oracle (oracle/), on the same seeded inputs,
  (2) the golden vectors produced by the unmodified reference (tests/golden/),
  (3) where available, the reference's own CUDA kernels compiled into oracle/_ref/ (bit-exact for aabb).
Tolerances are BASELINE.json's: SDF 1e-4 abs (we assert 2e-6), depth 1e-4 x scene scale on hits, normals 1e-3,
hit masks exact; gradients 1e-3 of max|grad| per tensor (we assert 2e-4)."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import nglod_oracle as O
from helpers import rand5_model, fit3_model, cl_flat, make_args, torus_sdf
from nglod_b200.lib.models import OctreeSDF as OctreeSDF_cls

pytestmark = pytest.mark.gpu
DEV = "cuda"

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))


def _ref(name):
    import build_ref
    return build_ref.load_ref(name)


# ------------------------------------------------------------------------------------------------ aabb
def _ray_sets():
    torch.manual_seed(2)
    o1, d1 = O.look_at([-2.8, 2.8, -2.8], [0, 0, 0], 1280, 720, fov=30.0)
    g = torch.Generator().manual_seed(9)
    o2 = torch.rand(100003, 3, generator=g) * 4 - 2                    # inside and outside origins, ragged n
    d2 = F.normalize(torch.randn(100003, 3, generator=g), dim=1)
    o3 = torch.tensor([[2.0, 0.0, 0.0], [0.0, -3.0, 0.0], [0.5, 0.5, 4.0], [1.0, 1.0, 1.0], [1.0, 0.0, 0.0],
                       [-2.0, 0.999, 0.0], [2.0, 2.0, 2.0]])
    d3 = torch.tensor([[-1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, -1.0], [-1.0, -1.0, -1.0], [-1.0, 0.0, 0.0],
                       [1.0, 0.0, 0.0], [0.0, -0.0, -1.0]])               # axis-aligned: zero components -> inf
    d3[3] = F.normalize(d3[3], dim=0)
    return [(o1, d1), (o2, d2), (o3, d3), (o2[:1], d2[:1]), (o2[:0], d2[:0])]


def test_aabb_bit_exact_vs_oracle_and_reference_kernel():
    from nglod_b200 import ops
    ref = _ref("ref_sol_nglod")
    for o, d in _ray_sets():
        x, t, hit = ops.aabb(o.to(DEV), d.to(DEV))
        ox, ot, oh = O.aabb(o, d)
        assert torch.equal(hit.cpu(), oh)
        assert torch.equal(t.cpu().view(torch.int32), ot.view(torch.int32))
        assert torch.equal(x.cpu().view(torch.int32), ox.view(torch.int32))
        if ref is not None and o.shape[0] > 0:
            rx, rt, rh = ref.aabb(o.to(DEV).contiguous(), d.to(DEV).contiguous())
            assert torch.equal(hit, rh)
            assert torch.equal(t.view(torch.int32), rt.view(torch.int32))
            assert torch.equal(x.view(torch.int32), rx.view(torch.int32))


def test_reference_kernel_available():
    """Not a failure if absent, but say so loudly: without oracle/_ref the aabb / mesh2sdf oracles are unpinned."""
    if _ref("ref_sol_nglod") is None or _ref("ref_mesh2sdf") is None:
        pytest.skip("oracle/_ref not built (python oracle/build_ref.py) -- native oracles unpinned on this box")


# ------------------------------------------------------------------------------------------------ sdf forward
def test_sdf_forward_vs_golden_and_oracle(rand5, fit3):
    net, _ = rand5_model(DEV)
    net.math_mode = "fp32"                      # the CUDA-core correctness anchor (tensor-core path: see below)
    x = torch.from_numpy(rand5["x"]).to(DEV)
    net.eval()
    with torch.no_grad():
        for l in range(5):
            d = net.sdf(x, lod=l)
            assert d.shape == (x.shape[0], 1)
            assert np.abs(d.cpu().numpy() - rand5[f"sdf_lod{l}"]).max() < 2e-6
        lst = net.sdf(x, return_lst=True)
        assert np.abs(torch.stack(lst).cpu().numpy() - rand5["sdf_lst"]).max() < 2e-6
        net.lod = 3
        assert np.abs(net(x).cpu().numpy() - rand5["forward_lod3"]).max() < 2e-6
    net3, _ = fit3_model(fit3, DEV)
    net3.math_mode = "fp32"
    x3 = torch.from_numpy(fit3["x"]).to(DEV)
    with torch.no_grad():
        for l in range(3):
            assert np.abs(net3.sdf(x3, lod=l).cpu().numpy() - fit3[f"sdf_lod{l}"]).max() < 2e-6


@pytest.mark.parametrize("n", [1, 31, 32, 33, 255, 1000, 65537])
def test_sdf_forward_ragged_sizes(n):
    net, _ = rand5_model(DEV)
    onet = O.OracleNet(net.state_dict())
    g = torch.Generator().manual_seed(n)
    x = torch.rand(n, 3, generator=g) * 2.4 - 1.2
    with torch.no_grad():
        ref = onet.sdf(x, lod=4)
        for mode in ("fp32", "tc"):
            net.math_mode = mode
            assert (net.sdf(x.to(DEV), lod=4).cpu() - ref).abs().max() < 2e-6


def test_sdf_empty_and_batched_shapes():
    net, _ = rand5_model(DEV)
    with torch.no_grad():
        assert net.sdf(torch.zeros(0, 3, device=DEV), lod=2).shape == (0, 1)
        x = torch.rand(7, 5, 3, device=DEV)
        d = net.sdf(x, lod=2)
        assert d.shape == (7, 5, 1)
        assert torch.allclose(d.reshape(-1, 1), net.sdf(x.reshape(-1, 3), lod=2))


def test_feature_volume_forward_vs_grid_sample():
    net, _ = rand5_model(DEV)
    onet = O.OracleNet(net.state_dict())
    g = torch.Generator().manual_seed(4)
    x = torch.rand(5000, 3, generator=g) * 2.2 - 1.1
    for i in (0, 3, 4):
        got = net.features[i](x.to(DEV)).cpu()
        assert (got - onet.sample(i, x)).abs().max() < 1e-7


# ------------------------------------------------------------------------------------------------ backward
def _check_grads(net, gold, tag, lods, tol=2e-4):
    for i in range(max(lods) + 1):
        g = cl_flat(net.features[i].fm.grad).cpu().numpy()
        if i <= 1:
            ref = gold[f"{tag}_fm{i}"]
            scale = np.abs(ref).max()
            assert np.abs(g - ref).max() / scale < tol
        else:
            ref = gold[f"{tag}_fm{i}_val"]
            scale = max(np.abs(ref).max(), 1e-12)
            assert np.abs(g[gold[f"{tag}_fm{i}_idx"]] - ref).max() / scale < tol
        s = gold[f"{tag}_fm{i}_sum"]
        assert abs(float(g.astype(np.float64).sum()) - s[0]) < tol * s[1]
    for i in range(max(lods) + 1, 5):
        assert net.features[i].fm.grad is None
    for l in range(5):
        seq = net.louts[l]
        if l in lods:
            for k, p in (("0.weight", seq[0].weight), ("0.bias", seq[0].bias), ("2.weight", seq[2].weight),
                         ("2.bias", seq[2].bias)):
                ref = gold[f"{tag}_louts{l}.{k}"]
                assert np.abs(p.grad.cpu().numpy() - ref).max() / np.abs(ref).max() < tol
        else:
            assert seq[0].weight.grad is None


@pytest.mark.parametrize("sum_lods", [True, False])
def test_autograd_backward_vs_golden(rand5, sum_lods):
    """sum_lods=True: forward from the prefix-summed grid, 8-corner scatter into its gradient, dense restriction down
    the LOD chain; False: the per-LOD gather / scatter.  Same golden gradients, same tolerance."""
    x = torch.from_numpy(rand5["x"]).to(DEV)
    gt = torch.from_numpy(rand5["gt"]).to(DEV)
    for tag, lods in (("g4", [4]), ("g13", [1, 3])):
        net, _ = rand5_model(DEV)
        net.sum_lods = sum_lods
        assert (net.net_view(inference=False).summed is not None) == sum_lods
        loss = 0
        for l in lods:
            loss = loss + ((net.sdf(x, lod=l) - gt) ** 2).sum()
        loss = loss / x.shape[0]
        loss.backward()
        assert abs(loss.item() - float(rand5[f"{tag}_loss"])) < 1e-6
        _check_grads(net, rand5, tag, lods)
        if sum_lods:        # the scratch of the single-grid backward is handed back zeroed
            assert all(float(t.abs().max()) == 0.0 for t in net.summed_grad_scratch())


def test_small_grid_private_scatter_copies(rand5):
    """Heads of the 4^3 / 8^3 levels: with nglod_net_grad_t.scatter_scratch the CTAs scatter into private copies that a fold
    kernel sums -- same gradients as the single-copy scatter (fp32 summation order aside), scratch handed back zeroed."""
    from nglod_b200 import ops
    net, _ = rand5_model(DEV)
    g = torch.Generator(device=DEV).manual_seed(5)
    n = 120000
    x = torch.rand(n, 3, device=DEV, generator=g) * 2 - 1
    go = torch.randn(n, device=DEV, generator=g) / n
    view = net.net_view(inference=False)
    assert view.summed is not None
    ss = net.scatter_scratch()
    for lod in (0, 1):
        res = []
        for scratch in (None, ss):
            gg = [torch.zeros_like(f.fm.data, memory_format=torch.preserve_format) for f in net.features]
            dg = tuple(torch.zeros_like(p) for p in net.decoder_params(lod))
            ops.sdf_backward(view, lod, x, go, gg, dg, summed_scratch=net.summed_grad_scratch(), scatter_scratch=scratch)
            res.append((gg, dg))
        for i in range(lod + 1):
            a, b = res[0][0][i], res[1][0][i]
            assert float(a.abs().max()) > 0
            assert float((a - b).abs().max()) <= 1e-4 * float(a.abs().max())     # summation order (atomics) differs
        for a, b in zip(res[0][1], res[1][1]):
            assert float((a - b).abs().max()) <= 1e-4 * float(a.abs().max())
        assert float(ss.abs().max()) == 0.0
        assert all(float(t.abs().max()) == 0.0 for t in net.summed_grad_scratch())


def test_grad_x_autodiff_and_finitediff_vs_golden(rand5):
    from nglod_b200.lib.diffutils import gradient
    net, _ = rand5_model(DEV)
    net.lod = 4
    x = torch.from_numpy(rand5["x"][:512]).to(DEV)
    ga = gradient(x.clone(), net, method="autodiff").cpu().numpy()
    assert np.abs(ga - rand5["autodiff_lod4"]).max() < 1e-5
    gf = gradient(x.clone(), net, method="finitediff").cpu().numpy()
    assert np.abs(gf - rand5["finitediff_lod4"]).max() < 1e-4


@pytest.mark.parametrize("sum_lods", [True, False])
def test_fused_train_step_vs_oracle_adam(rand5, sum_lods):
    """FusedTrainer (flat buffers, fused fwd+loss+bwd per LOD, Adam kernel) vs torch autograd + torch.optim.Adam
    on the oracle, two steps, all five heads in the loss."""
    from nglod_b200.lib.trainer import FusedTrainer
    net, _ = rand5_model(DEV)
    net.sum_lods = sum_lods
    onet = O.OracleNet(net.state_dict(), requires_grad=True)
    opt = torch.optim.Adam(onet.parameters(), lr=1e-3)
    tr = FusedTrainer(net, lr=1e-3)
    x = torch.from_numpy(rand5["x"])
    gt = torch.from_numpy(rand5["gt"])
    for step in range(2):
        loss_ref = O.l2_loss_and_grads(onet, x, gt, [0, 1, 2, 3, 4])
        loss = tr.step(x.to(DEV), gt.to(DEV))
        assert abs(loss.item() - loss_ref.item()) < 1e-5 * max(1.0, abs(loss_ref.item()))
        if step == 0:
            for i in range(5):
                g = cl_flat(net.features[i].fm.grad).cpu()
                r = cl_flat(onet.fm[i].grad)
                assert (g - r).abs().max() / r.abs().max() < 2e-4
        opt.step()
    # after two Adam steps parameters agree (Adam normalises the step to ~lr, so compare against lr)
    for i in range(5):
        a = net.features[i].fm.detach().cpu()
        b = onet.fm[i].detach()
        assert (a - b).abs().max() < 2e-4
    w = net.louts[4][0].weight.detach().cpu()
    assert (w - onet.dec[4][0].detach()).abs().max() < 2e-4


# ------------------------------------------------------------------------------------------------ tracer
def _check_trace(rb, gold, prefix, conv=None):
    hit = gold[prefix + "_hit"]
    got_hit = rb.hit.cpu().numpy()
    assert np.array_equal(got_hit, hit), f"hit mask differs on {(got_hit != hit).sum()} rays"
    depth = np.abs(rb.depth.cpu().numpy() - gold[prefix + "_depth"])[:, 0]
    xerr = np.abs(rb.x.cpu().numpy() - gold[prefix + "_x"]).max(axis=1)
    nerr = np.abs(rb.normal.cpu().numpy() - gold[prefix + "_normal"]).max(axis=1)
    sel = hit if conv is None else conv
    if sel.any():
        assert depth[sel].max() < 2e-4          # 1e-4 x scene scale (box side 2)
        # x lags depth by one march step (SphereTracer.py:102-114), so a ray whose last |d| sits within float noise
        # of min_dis may stop one step apart: x then differs by that step (<= min_dis = 3e-4), depth by much less.
        assert xerr[sel].max() < 3e-4 + 1.5e-4
        assert (xerr[sel] > 1e-4).mean() < 0.01
        assert nerr[sel].max() < 1e-3 or (nerr[sel] > 1e-3).mean() < 0.005
        print(f"{prefix}: {int(sel.sum())} rays checked; depth max {depth[sel].max():.2e}; x max {xerr[sel].max():.2e} "
              f"({int((xerr[sel] > 1e-4).sum())} > 1e-4); normal max {nerr[sel].max():.2e} "
              f"({int((nerr[sel] > 1e-3).sum())} > 1e-3)")
    assert nerr[~hit].max(initial=0) == 0.0     # normals are exactly zero where not hit
    return depth, nerr


@pytest.mark.parametrize("math_mode", ["fp32", "tc"])
def test_sphere_tracer_vs_golden(rand5, fit3, math_mode):
    from nglod_b200.lib.tracer import SphereTracer
    net, args = rand5_model(DEV)
    net.math_mode = math_mode
    net.lod = 4
    tracer = SphereTracer(args)
    rb = tracer(net, torch.from_numpy(rand5["t1_ray_o"]).to(DEV), torch.from_numpy(rand5["t1_ray_d"]).to(DEV))
    _check_trace(rb, rand5, "t1")
    rb = SphereTracer(args, num_steps=12)(net, torch.from_numpy(rand5["t2_ray_o"]).to(DEV),
                                          torch.from_numpy(rand5["t2_ray_d"]).to(DEV))
    _check_trace(rb, rand5, "t2")
    net3, args3 = fit3_model(fit3, DEV)
    net3.math_mode = math_mode
    net3.lod = 2
    rb = SphereTracer(args3)(net3, torch.from_numpy(fit3["t1_ray_o"]).to(DEV), torch.from_numpy(fit3["t1_ray_d"]).to(DEV))
    depth, nerr = _check_trace(rb, fit3, "t1", conv=fit3["t1_converged"])
    # non-converged "hits" (budget exhausted / oscillation stop) are chaotic in the reference itself: report only
    hit = fit3["t1_hit"]
    print(f"fit3 trace: {hit.sum()} hits, {fit3['t1_converged'].sum()} converged; "
          f"all-hit depth max {depth[hit].max():.2e}, normal max {nerr[hit].max():.2e}")


@pytest.mark.parametrize("math_mode", ["fp32", "tc"])
def test_sphere_tracer_vs_batch_loop_720p(fit3, math_mode):
    """Full 1280x720 frame: the persistent kernel against the reference's batch loop expressed with torch ops on
    the device, with the SAME sdf kernel as `net`.  The only arithmetic difference is x = o + d*t: torch's CUDA
    addcmul contracts it into an fma, the kernel follows the CPU reference (rounded product, then add; the golden
    vectors are CPU), so positions differ by <= 1 ulp and the iterated march by float noise."""
    from nglod_b200.lib.tracer import SphereTracer
    net3, args3 = fit3_model(fit3, DEV)
    net3.math_mode = math_mode
    net3.lod = 2
    torch.manual_seed(5)
    o, d = O.look_at([-2.8, 2.8, -2.8], [0, 0, 0], 1280, 720, fov=30.0)
    o, d = o.to(DEV), d.to(DEV)
    tr = SphereTracer(args3)
    rb = tr(net3, o, d)
    rg = tr._forward_generic(net3, o, d, track_min=False)
    mism = int((rb.hit != rg.hit).sum())
    conv = (net3(rg.x).abs() < 0.0003)[:, 0] & rg.hit & rb.hit
    dd = (rb.depth - rg.depth).abs()[:, 0]
    nn = (rb.normal - rg.normal).abs().max(dim=1)[0]
    print(f"720p: {int(rg.hit.sum())} hits, {mism} hit-mask mismatches, {int(conv.sum())} converged; "
          f"depth diff max {float(dd[conv].max()):.2e} (>1e-4: {int((dd[conv] > 1e-4).sum())}); "
          f"normal diff max {float(nn[conv].max()):.2e} (>1e-3: {int((nn[conv] > 1e-3).sum())})")
    assert mism <= 2                                   # hit masks: exact up to threshold-straddling rays
    assert float((dd[conv] > 2e-4).float().mean()) < 1e-3
    assert float((nn[conv] > 1e-3).float().mean()) < 5e-3
    h = rb.hit
    assert h.sum() > 50000
    assert (rb.x[h].abs() <= 1.0).all()
    assert ((rb.normal[h].norm(dim=1) - 1).abs() < 1e-4).all()
    assert (rb.normal[~h] == 0).all()


def test_tracer_edge_cases():
    from nglod_b200.lib.tracer import SphereTracer
    net, args = rand5_model(DEV)
    net.lod = 4
    e = torch.zeros(0, 3, device=DEV)
    rb = SphereTracer(args)(net, e, e)
    assert rb.hit.shape == (0,) and rb.x.shape == (0, 3)
    # num_steps = 1 and a far plane nothing can pass
    o = torch.tensor([[0.0, 0.0, -3.0]], device=DEV).repeat(64, 1)
    d = torch.tensor([[0.0, 0.0, 1.0]], device=DEV).repeat(64, 1)
    rb = SphereTracer(args, num_steps=1, camera_clamp=[0, 1.0])(net, o, d)
    assert not rb.hit.any()                         # |t| = 2 >= far -> flag false


# ------------------------------------------------------------------------------------------------ mesh2sdf
def test_mesh2sdf_vs_reference_kernel_and_oracle():
    from nglod_b200 import ops
    from nglod_b200.lib.torchgp import icosphere, torus
    ref = _ref("ref_mesh2sdf")
    g = torch.Generator().manual_seed(21)
    for name, (V, Fc) in (("ico3", icosphere(3)), ("torus", torus(nu=48, nv=24))):
        mesh = V[Fc].to(DEV).contiguous()
        pts = (torch.rand(20011, 3, generator=g) * 2 - 1).to(DEV)
        d = ops.mesh2sdf_gpu(pts, mesh)[0]
        sub = slice(0, 1500)
        do = O.mesh2sdf(pts[sub].cpu(), mesh.cpu())
        assert (d[sub].cpu() - do).abs().max() < 1e-6
        assert torch.equal(d[sub].cpu() < 0, do < 0)
        if ref is not None:
            dr = ref.mesh2sdf_gpu(pts, mesh)[0]
            assert (d - dr).abs().max() < 1e-6, name
            assert int(((d < 0) != (dr < 0)).sum()) == 0, name
    # analytic check on the sphere
    V, Fc = icosphere(4)
    pts = (torch.rand(50000, 3, generator=g) * 2 - 1).to(DEV)
    d = ops.mesh2sdf_gpu(pts, V[Fc].to(DEV))[0]
    r = pts.norm(dim=1)
    assert ((d - (r - 1)).abs() < 0.01).all()
    far = (r - 1).abs() > 0.01
    assert torch.equal((d < 0)[far], (r < 1)[far])
    assert ops.mesh2sdf_gpu(pts[:0], V[Fc].to(DEV))[0].shape == (0,)


def test_mesh2sdf_hierarchy_is_bit_identical():
    """The Morton-sorted patches + per-warp patch / triangle culls only skip pairs the exact tests dismiss anyway:
    every distance and sign equals the brute-force walk bit for bit (shuffled, duplicated and degenerate triangles, points
    outside the box, all batch-size regimes: unsorted / sorted / sliced)."""
    import os
    from nglod_b200 import ops
    from nglod_b200.lib.torchgp import icosphere, torus, normalize, point_sample
    g = torch.Generator().manual_seed(5)
    V, Fc = normalize(*[t.to(DEV) for t in torus(0.6, 0.25, 96, 48)])
    tri = V[Fc].contiguous()
    extra = torch.cat([tri[:64], tri[:32, :1].expand(-1, 3, -1)], 0)          # duplicates + zero-area triangles
    tri = torch.cat([tri, extra], 0)[torch.randperm(tri.shape[0] + 96, generator=g).to(DEV)].contiguous()
    near = point_sample(V, Fc, ["near", "trace"], 60000)
    box = (torch.rand(80000, 3, generator=g) * 2.4 - 1.2).to(DEV)
    pts = torch.cat([near, box], 0)[torch.randperm(200000, generator=g).to(DEV)].contiguous()

    def run(p, cull):
        return ops.mesh2sdf_gpu(p, tri, force_walk=not cull)[0]

    for n in (200000, 30000, 3000, 37):
        p = pts[:n].contiguous()
        a, b = run(p, True), run(p, False)
        assert torch.equal(a.view(torch.int32), b.view(torch.int32)), n
        assert int((a < 0).sum()) > 0 or n < 100
    # a mesh too small for tiles takes the plain walk
    Vs, Fs = icosphere(1)
    ts = Vs[Fs].to(DEV).contiguous()
    d = ops.mesh2sdf_gpu(pts[:5000].contiguous(), ts)[0]
    assert torch.isfinite(d).all()


def _cube_mesh(k=8):
    """Axis-aligned cube [-0.5, 0.5]^3, k x k quads per face (12 k^2 triangles)."""
    g = torch.linspace(-0.5, 0.5, k + 1)
    u, v = torch.meshgrid(g, g, indexing="ij")
    tris = []
    for axis in range(3):
        for side in (-0.5, 0.5):
            P = torch.zeros(k + 1, k + 1, 3)
            P[..., axis] = side
            P[..., (axis + 1) % 3] = u
            P[..., (axis + 2) % 3] = v
            a, b, c, d = P[:-1, :-1], P[1:, :-1], P[1:, 1:], P[:-1, 1:]
            tris.append(torch.stack([a, b, c], -2).reshape(-1, 3, 3))
            tris.append(torch.stack([a, c, d], -2).reshape(-1, 3, 3))
    return torch.cat(tris, 0)


def test_mesh2sdf_hierarchy_grazing_and_grid_sizes():
    """The sign half bins points under each triangle's projection widened by the rounding error of the exact test, which
    blows up when a stab line grazes the triangle's plane: cubes whose faces are parallel to three stab directions
    (determinant exactly 0: direction skipped) and rotated off them by 1e-6 .. 1e-2 rad (determinants down to the 1e-8
    threshold) must still match the brute-force walk bit for bit.  Also every projected-grid size (32 .. 512 cells a side)."""
    import math
    import os
    from nglod_b200 import ops
    from nglod_b200.lib.torchgp import icosphere
    g = torch.Generator().manual_seed(9)

    def both(p, tri):
        return ops.mesh2sdf_gpu(p, tri)[0], ops.mesh2sdf_gpu(p, tri, force_walk=True)[0]

    cube = _cube_mesh(8)
    pts = (torch.rand(40000, 3, generator=g) * 2 - 1).to(DEV)
    # points exactly on the planes / lines the cube's faces and edges span, where the barycentric tests sit on their limits
    snap = pts[:8000].clone()
    snap[:4000, 0] = 0.5
    snap[4000:, 1] = (snap[4000:, 1] * 8).round() / 8
    pts = torch.cat([pts, snap], 0).contiguous()
    for angle in (0.0, 1e-6, 1e-4, 1e-2, 0.3):
        ax = torch.tensor([0.3, 0.5, 0.81]); ax = ax / ax.norm()
        K = torch.tensor([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        R = torch.eye(3) + math.sin(angle) * K + (1 - math.cos(angle)) * (K @ K)
        tri = (cube @ R.T).to(DEV).contiguous()
        a, b = both(pts, tri)
        assert torch.equal(a.view(torch.int32), b.view(torch.int32)), angle
        if angle == 0.3:
            inside = (pts @ R.to(DEV)).abs().max(dim=1).values < 0.5 - 1e-4       # rotate back: inside the cube
            assert (a[inside] < 0).all()
    # a scene that was not normalised into the unit cube (x5.3, off-centre): the projected grids and the rounding margin
    # follow the extent of the call; the spatial sorts clamp, so this is slow but must stay exact
    V5, F5 = icosphere(3)
    a, b = both((pts[:20000] * 5.3 + 0.7).contiguous(), ((V5[F5] * 5.3 * 0.8) + 0.7).to(DEV).contiguous())
    assert torch.equal(a.view(torch.int32), b.view(torch.int32)) and int((a < 0).sum()) > 1000
    a, b = both(pts, (_cube_mesh(4) * 1.3).to(DEV).contiguous())                 # 192 triangles: the smallest grid, G = 32
    assert torch.equal(a.view(torch.int32), b.view(torch.int32))
    for sub, n in ((2, 20000), (3, 20000), (4, 30000), (6, 50000)):              # 320 .. 81 920 triangles: G = 64 .. 512
        V, Fc = icosphere(sub)
        tri = V[Fc].to(DEV).contiguous()
        p = (torch.rand(n, 3, generator=g) * 2.2 - 1.1).to(DEV)
        a, b = both(p, tri)
        assert torch.equal(a.view(torch.int32), b.view(torch.int32)), sub


def test_sample_mesh_kernel_distribution():
    """nglod_sample_mesh vs the reference recipe (torchgp/*.py): parity is distributional (RNG streams differ), so check the
    moments the recipe fixes -- face frequency ~ area, uniform barycentric density, noise std, cube range -- and the exact
    things: technique-major order, points on their face, determinism by seed."""
    from nglod_b200 import ops, _lib
    from nglod_b200.lib.torchgp import torus, normalize, per_face_normals
    V, Fc = normalize(*[t.to(DEV) for t in torus(0.6, 0.25, 24, 12)])
    # make the areas very unequal and add a zero-area face that must never be drawn
    Fc = torch.cat([Fc, Fc[:1, [0, 0, 1]]], 0)
    cdf = ops.mesh_area_cdf(V, Fc)
    tri = V[Fc]
    area = 0.5 * torch.linalg.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]).norm(dim=1)
    assert torch.allclose(cdf, torch.cumsum(area.double(), 0).float(), rtol=1e-5, atol=1e-7)
    assert (cdf[1:] >= cdf[:-1]).all()
    S = 400000
    pts, faces = ops.sample_mesh(V, Fc, cdf, ["rand", "near", "trace"], S, variance=0.01, seed=1234, return_faces=True)
    assert pts.shape == (3 * S, 3) and faces.shape == (3 * S,)
    rand, near, trace = pts[:S], pts[S:2 * S], pts[2 * S:]
    # rand: U[-1,1)^3
    assert (faces[:S] == -1).all() and rand.min() >= -1 and rand.max() < 1
    assert rand.mean(0).abs().max() < 0.01 and ((rand.var(0) - 1.0 / 3).abs() < 0.01).all()
    # trace: on the drawn face (barycentric coordinates in [0,1], zero plane distance), faces ~ area
    ft = faces[2 * S:].long()
    assert ft.min() >= 0 and ft.max() < Fc.shape[0] - 1                         # the zero-area face never comes up
    a, b, c = tri[ft, 0], tri[ft, 1], tri[ft, 2]
    n = torch.nn.functional.normalize(torch.linalg.cross(b - a, c - a), dim=1)
    assert ((trace - a) * n).sum(1).abs().max() < 1e-6
    freq = torch.bincount(ft, minlength=Fc.shape[0]).double() / S
    p = (area / area.sum()).double()
    sigma = torch.sqrt(p * (1 - p) / S) + 1e-9
    assert ((freq - p).abs() / sigma).max() < 6.0                              # every face within 6 sigma of its area share
    # uniform on the triangle: mean of the barycentric weights is 1/3 each
    e1, e2, dd = b - a, c - a, trace - a
    d11, d12, d22, r1, r2 = (e1 * e1).sum(1), (e1 * e2).sum(1), (e2 * e2).sum(1), (dd * e1).sum(1), (dd * e2).sum(1)
    det = d11 * d22 - d12 * d12
    w12 = torch.stack([(d22 * r1 - d12 * r2) / det, (d11 * r2 - d12 * r1) / det], 1)
    assert (w12.mean(0) - 1.0 / 3).abs().max() < 3e-3 and w12.min() > -1e-4 and w12.sum(1).max() < 1 + 1e-4
    # near: surface point + N(0, 0.01^2) per coordinate, isotropic and uncorrelated
    fn = faces[S:2 * S].long()
    an, nn = tri[fn, 0], torch.nn.functional.normalize(per_face_normals(V, Fc)[fn], dim=1)
    off = ((near - an) * nn).sum(1)                                             # normal component of the noise
    assert abs(off.mean().item()) < 1e-4 and abs(off.std().item() - 0.01) < 2e-4
    kurt = ((off / off.std()) ** 4).mean().item()
    assert abs(kurt - 3.0) < 0.1                                                # gaussian, not uniform
    # determinism / seeds
    again = ops.sample_mesh(V, Fc, cdf, ["rand", "near", "trace"], S, variance=0.01, seed=1234)
    assert torch.equal(again, pts)
    other = ops.sample_mesh(V, Fc, cdf, ["rand", "near", "trace"], S, variance=0.01, seed=1235)
    assert not torch.equal(other, pts)
    # host wrappers: manual_seed reproduces, sample_surface returns the faces' (unnormalised) normals
    from nglod_b200.lib.torchgp import point_sample, sample_surface
    torch.manual_seed(3); p1 = point_sample(V, Fc, ["rand", "near", "near", "trace", "trace"], 1000)
    torch.manual_seed(3); p2 = point_sample(V, Fc, ["rand", "near", "near", "trace", "trace"], 1000)
    assert p1.shape == (5000, 3) and torch.equal(p1, p2)
    assert not torch.equal(p1[1000:2000], p1[2000:3000])                        # the two 'near' blocks are different draws
    sp, sn = sample_surface(V, Fc, 2000)
    assert sp.shape == (2000, 3) and sn.shape == (2000, 3)
    assert point_sample(V, Fc, ["rand"], 10).shape == (10, 3) and point_sample(V, Fc, [], 10).shape == (0, 3)
    lib = _lib.load()
    assert lib.nglod_sample_mesh(None, None, 0, None, None, 40, 1, 0.0, 0, None, None, None) == _lib.EINVAL


def test_sample_mesh_stream_is_philox():
    """The sampler's random stream is Philox-4x32-10 with counter = sample index, key = seed: 'rand' samples equal the
    oracle's restatement of the published algorithm exactly (which test_philox_known_answers pins to Random123's vectors)."""
    from nglod_b200 import ops
    V = torch.zeros(3, 3, device=DEV)
    Fc = torch.tensor([[0, 1, 2]], device=DEV)
    for seed in (0, 1, 0x299f31d0a4093822, 2 ** 62 - 1):
        pts = ops.sample_mesh(V, Fc, None, ["rand"], 5000, seed=seed).cpu()
        for i in (0, 1, 2, 31, 32, 255, 4999):
            want = torch.tensor(O.sample_uniform_philox(seed, i), dtype=torch.float32)
            assert torch.equal(pts[i], want), (seed, i)


def test_mesh_dataset_protocol():
    from nglod_b200.lib.datasets import MeshDataset
    from nglod_b200.lib.torchgp import torus
    ds = MeshDataset(make_args(["--num-samples", "2000"]), mesh=torus(nu=32, nv=16))
    assert len(ds) == 5 * 2000 and ds.num_shapes() == 1
    p, d = ds[3]
    assert p.shape == (3,) and d.shape == (1,)
    assert float(ds.V.norm(dim=1).max()) == pytest.approx(1.0, abs=1e-6)
    assert (ds.pts[:2000].abs() <= 1).all()                      # 'rand' block
    assert ds.d[6000:].abs().max() < 1e-4                        # 'trace' samples lie on the surface
    assert 0.002 < ds.d[2000:6000].abs().mean() < 0.02           # 'near' samples: N(0, 0.01) off the surface


# ------------------------------------------------------------------------------------------------ renderer
def test_renderer_render_with_ao_vs_golden(fit3):
    from nglod_b200.lib.tracer import SphereTracer
    from nglod_b200.lib.renderer import Renderer
    net3, _ = fit3_model(fit3, DEV)
    net3.lod = 2
    rargs = make_args(["--num-lods", "3", "--render-res", "96", "54", "--ao"])
    r = Renderer(SphereTracer(rargs), args=rargs, device=DEV)
    rb = r.render(net3, torch.from_numpy(fit3["t1_ray_o"]).to(DEV), torch.from_numpy(fit3["t1_ray_d"]).to(DEV))
    assert rb.hit.shape == (96, 54, 1) and rb.normal.shape == (96, 54, 3)
    assert np.array_equal(rb.hit.cpu().numpy(), fit3["r1_hit"])
    conv = fit3["t1_converged"].reshape(96, 54)
    assert np.abs(rb.ao.cpu().numpy() - fit3["r1_ao"])[conv].max() < 2e-3
    assert np.abs(rb.relative_depth.cpu().numpy() - fit3["r1_relative_depth"])[conv].max() < 2e-5


def test_shade_images_layout_and_shadow(fit3):
    from nglod_b200.lib.tracer import SphereTracer
    from nglod_b200.lib.renderer import Renderer
    net3, _ = fit3_model(fit3, DEV)
    net3.lod = 2
    rargs = make_args(["--num-lods", "3", "--render-res", "160", "90", "--shadow", "--ground-height", "-0.3", "--ao"])
    r = Renderer(SphereTracer(rargs), args=rargs, device=DEV)
    out = r.shade_images(net3, f=rargs.camera_origin, t=rargs.camera_lookat, fov=rargs.camera_fov)
    assert out.rgb.shape == (90, 160, 3) and out.hit.shape == (90, 160, 1) and not out.rgb.is_cuda
    assert 0.0 <= float(out.rgb.min()) and float(out.rgb.max()) <= 1.0 + 1e-5
    assert out.shadow.any() and out.hit.any()
    img = out.image().byte().numpy()
    assert img.rgb.dtype == np.uint8 and img.rgb.shape == (90, 160, 3)


# ------------------------------------------------------------------------------------------------ errors
def test_fails_loudly_instead_of_falling_back():
    from nglod_b200 import ops, _lib
    from nglod_b200.lib.models import OctreeSDF
    with pytest.raises(RuntimeError):
        ops.aabb(torch.zeros(4, 3), torch.zeros(4, 3))             # CPU tensors: no CPU path
    net = OctreeSDF(make_args(["--num-lods", "2"]))               # parameters on the CPU
    with pytest.raises(RuntimeError):
        net.sdf(torch.zeros(4, 3, device=DEV), lod=1)
    # the library itself takes feature-dim 32 / hidden-dim 128 only (smaller models are zero-padded by the host class, see
    # test_model_variants_forward_backward_trace); anything else is refused, nothing is silently approximated
    with pytest.raises(RuntimeError, match="not supported"):
        OctreeSDF(make_args(["--num-lods", "2", "--hidden-dim", "256"]))
    net64 = OctreeSDF(make_args(["--num-lods", "2", "--hidden-dim", "64"])).to(DEV)
    raw = ops.NetView([f.fm.data for f in net64.features], [tuple(p.data for p in net64.decoder_params(i)) for i in range(2)])
    with pytest.raises(RuntimeError, match="unsupported"):
        ops.sdf_forward(raw, 1, torch.zeros(4, 3, device=DEV))
    lib = _lib.load()
    assert lib.nglod_aabb(None, None, 5, None, None, None, None) == _lib.EINVAL
    assert lib.nglod_mesh2sdf(None, -1, None, 0, None, None) == _lib.EINVAL
    assert lib.nglod_aabb(None, None, 0, None, None, None, None) == 0


# ------------------------------------------------------------------------------------------------ tensor-core path
def test_tc_gemm_selftest():
    """The tcgen05 plumbing in isolation: operand layout, descriptors, 3xTF32 sequence, TMEM read-back."""
    from nglod_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(3)
    A = torch.randn(128, 40, generator=g)
    B = torch.randn(128, 40, generator=g)
    A[:, 36:] = 0
    D = torch.full((128, 128), float("nan"), device=DEV)
    Ad, Bd = A.to(DEV), B.to(DEV)
    rc = lib.nglod_debug_tc_gemm(ctypes.c_void_p(Ad.data_ptr()), ctypes.c_void_p(Bd.data_ptr()),
                                 ctypes.c_void_p(D.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    torch.cuda.synchronize()
    ref = (A.double() @ B.double().T)
    err = (D.cpu().double() - ref).abs().max().item()
    print(f"tc gemm 3xTF32 max abs err {err:.3e} (|ref| max {ref.abs().max():.1f})")
    assert err < 2e-5            # single-pass TF32 would be ~1e-2 here


def test_sdf_forward_tensor_core_vs_golden(rand5, fit3):
    net, _ = rand5_model(DEV)
    net.math_mode = "tc"
    x = torch.from_numpy(rand5["x"]).to(DEV)
    with torch.no_grad():
        for l in range(5):
            err = np.abs(net.sdf(x, lod=l).cpu().numpy() - rand5[f"sdf_lod{l}"]).max()
            print(f"tc sdf lod{l} max err {err:.2e}")
            assert err < 5e-6
    net3, _ = fit3_model(fit3, DEV)
    net3.math_mode = "tc"
    x3 = torch.from_numpy(fit3["x"]).to(DEV)
    with torch.no_grad():
        for l in range(3):
            err = np.abs(net3.sdf(x3, lod=l).cpu().numpy() - fit3[f"sdf_lod{l}"]).max()
            print(f"tc sdf fit3 lod{l} max err {err:.2e}")
            assert err < 5e-6
        for n in (1, 127, 129, 1000, 100001):
            xx = torch.rand(n, 3, device=DEV) * 2.4 - 1.2
            net3.math_mode = "tc"
            a = net3.sdf(xx, lod=2)
            net3.math_mode = "fp32"
            b = net3.sdf(xx, lod=2)
            assert (a - b).abs().max() < 5e-6


# ------------------------------------------------------------------------------------------------ fp16 grid storage
def test_pack_grid_fp16_layout_bit_exact():
    """nglod_pack_grid_fp16: line (z,y,x0), chunk c = {x0 ch 4c..4c+3, x1 ch 4c..4c+3}, round-to-nearest-even."""
    from nglod_b200 import ops
    for R in (1, 4, 7, 16):
        S = R + 1
        g = torch.Generator().manual_seed(R)
        fm = (torch.randn(1, 32, S, S, S, generator=g) * 0.01).to(DEV).contiguous(memory_format=torch.channels_last_3d)
        got = ops.pack_grid_fp16(fm).cpu().view(torch.float16).reshape(S, S, R, 8, 2, 4)
        cl = fm[0].permute(1, 2, 3, 0).cpu()                                  # [z, y, x, 32]
        a = cl[:, :, :-1].half().reshape(S, S, R, 8, 4)
        b = cl[:, :, 1:].half().reshape(S, S, R, 8, 4)
        assert torch.equal(got[..., 0, :], a) and torch.equal(got[..., 1, :], b)


def test_summed_grid_identity(rand5):
    """The prefix-summed grid: (1) its nodes hold sum_l trilinear(grid_l, node) (vs the oracle's grid_sample at the node
    positions), (2) ONE trilinear sample of it equals the reference's running sum of per-LOD samples at arbitrary x,
    including clamped points outside the box, (3) grids that do not nest are refused."""
    from nglod_b200 import ops
    net, _ = rand5_model(DEV)
    onet = O.OracleNet(net.state_dict())
    grids = [f.fm.data for f in net.features]
    view = ops.NetView.grids_only(grids)
    g = torch.Generator().manual_seed(11)
    x = torch.cat([torch.rand(20000, 3, generator=g) * 2.4 - 1.2, torch.tensor([[1.0, 1.0, 1.0], [-1.0, 0.5, 1.0]])])
    for lod in (0, 2, 4):
        sg = ops.build_summed_grid(view, lod)
        S = grids[lod].shape[-1]
        lin = torch.arange(S, dtype=torch.float32) * (2.0 / (S - 1)) - 1.0
        zz, yy, xx = torch.meshgrid(lin, lin, lin, indexing="ij")
        nodes = torch.stack([xx, yy, zz], dim=-1).reshape(-1, 3)
        want = sum(onet.sample(i, nodes) for i in range(lod + 1))                      # [S^3, 32], z-major like the grid
        got = sg[0].permute(1, 2, 3, 0).reshape(-1, 32).cpu()
        assert (got - want).abs().max() < 2e-8
        ref = sum(onet.sample(i, x) for i in range(lod + 1))
        one = ops.sdf_features(ops.NetView.grids_only([sg]), 0, x.to(DEV)).cpu()
        err = (one - ref).abs().max().item()
        print(f"summed grid lod{lod}: node err {(got - want).abs().max():.1e}; sample-vs-running-sum err {err:.1e}")
        assert err < 3e-8
    # the model builds its levels in ascending order, each as the prolongation of the previous one + its own grid
    # (nglod_net_t.summed[lod-1] given): same grids as the direct builds
    for lod, sg in enumerate(net.net_view().summed):
        assert (sg - ops.build_summed_grid(view, lod)).abs().max() < 2e-8
    odd = [torch.zeros(1, 32, r + 1, r + 1, r + 1, device=DEV).contiguous(memory_format=torch.channels_last_3d) for r in (4, 6)]
    with pytest.raises(RuntimeError):
        ops.build_summed_grid(ops.NetView.grids_only(odd), 1)


@pytest.mark.parametrize("math_mode", ["fp32", "tc"])
def test_sdf_forward_without_lod_summing(rand5, math_mode):
    """sum_lods=False keeps the per-LOD gather (what training uses): same golden tolerances."""
    net, _ = rand5_model(DEV)
    net.math_mode, net.sum_lods = math_mode, False
    assert net.net_view().summed is None
    x = torch.from_numpy(rand5["x"]).to(DEV)
    with torch.no_grad():
        for l in range(5):
            assert np.abs(net.sdf(x, lod=l).cpu().numpy() - rand5[f"sdf_lod{l}"]).max() < 5e-6


def test_sdf_forward_fp16_storage(rand5, fit3):
    """grid_storage='fp16': (1) equals the oracle run on the model's own `summed_state_dict()` (the fp16-rounded summed
    grid in the reference's format; fp32 arithmetic), (2) stays within BASELINE's 1e-4 of the reference on the unrounded
    weights, (3) the derived grids are rebuilt when a grid changes."""
    for maker, gold, nl in ((lambda: rand5_model(DEV), rand5, 5), (lambda: fit3_model(fit3, DEV), fit3, 3)):
        net, _ = maker()
        net.math_mode, net.grid_storage = "tc", "fp16"
        assert net.net_view().summed_half is not None
        x = torch.from_numpy(gold["x"])
        edge = torch.tensor([[1.0, 1.0, 1.0], [-1.0, -1.0, -1.0], [1.0, -0.3, 0.2], [0.1, 1.0, -1.0], [1.5, 0.0, -2.0]])
        xx = torch.cat([x, edge])
        with torch.no_grad():
            for l in range(nl):
                d = net.sdf(xx.to(DEV), lod=l).cpu()
                err_q = (d - O.OracleNet(net.summed_state_dict(l)).sdf(xx, lod=l)).abs().max().item()
                err_ref = np.abs(d[:x.shape[0]].numpy() - gold[f"sdf_lod{l}"]).max()
                print(f"fp16 storage lod{l}: vs oracle(summed, fp16-rounded) {err_q:.2e}; vs reference(fp32 weights) {err_ref:.2e}")
                assert err_q < 5e-6
                assert err_ref < 1e-4
            # staleness: an in-place update of the Parameter is seen through torch's version counter; writes behind
            # it (.data / raw pointers, e.g. the fused Adam kernel) need mark_grids_dirty()
            before = net.sdf(x.to(DEV), lod=nl - 1)
            net.features[nl - 1].fm.mul_(2.0)
            after = net.sdf(x.to(DEV), lod=nl - 1)
            ref2 = O.OracleNet(net.summed_state_dict()).sdf(x, lod=nl - 1)
            assert (after.cpu() - ref2).abs().max() < 5e-6 and (after - before).abs().max() > 1e-5
            net.features[0].fm.data.mul_(0.5)
            net.mark_grids_dirty()
            ref3 = O.OracleNet(net.summed_state_dict()).sdf(x, lod=nl - 1)
            assert (net.sdf(x.to(DEV), lod=nl - 1).cpu() - ref3).abs().max() < 5e-6
            # and the default fp32 storage follows the same updates, against the oracle on the ORIGINAL parameters
            net.grid_storage = "fp32"
            assert (net.sdf(x.to(DEV), lod=nl - 1).cpu() - O.OracleNet(net.state_dict()).sdf(x, lod=nl - 1)).abs().max() < 5e-6


def test_sphere_tracer_fp16_storage_vs_oracle_on_quantized_grids(rand5, fit3):
    from nglod_b200.lib.tracer import SphereTracer
    net3, args3 = fit3_model(fit3, DEV)
    net3.math_mode, net3.grid_storage, net3.lod = "tc", "fp16", 2
    o, d = torch.from_numpy(fit3["t1_ray_o"]), torch.from_numpy(fit3["t1_ray_d"])
    rb = SphereTracer(args3)(net3, o.to(DEV), d.to(DEV))
    onet = O.OracleNet(net3.summed_state_dict(2))
    onet.lod = 2
    res = O.sphere_trace(onet, o, d)
    hit = res["hit"].numpy()
    got = rb.hit.cpu().numpy()
    mism = int((got != hit).sum())
    conv = (onet(res["x"]).abs() < 0.0003)[:, 0].numpy() & hit & got
    dd = (rb.depth.cpu() - res["depth"]).abs()[:, 0].numpy()
    nn = (rb.normal.cpu() - res["normal"]).abs().max(dim=1)[0].numpy()
    print(f"fp16 tracer vs oracle(quantized): {int(hit.sum())} hits, {mism} mask mismatches, {int(conv.sum())} converged, "
          f"depth max {dd[conv].max():.2e}, normal max {nn[conv].max():.2e} (>1e-3: {int((nn[conv] > 1e-3).sum())})")
    assert mism <= 1
    assert dd[conv].max() < 2e-4
    assert (nn[conv] > 1e-3).mean() < 0.005
    # and against the reference on the UNROUNDED weights (golden): BASELINE tolerances on converged hits
    gh = fit3["t1_hit"]
    both = fit3["t1_converged"] & got & gh
    dg = np.abs(rb.depth.cpu().numpy() - fit3["t1_depth"])[:, 0]
    ng = np.abs(rb.normal.cpu().numpy() - fit3["t1_normal"]).max(axis=1)
    print(f"fp16 tracer vs reference(fp32 weights): mask mismatches {int((got != gh).sum())}, depth max {dg[both].max():.2e}, "
          f"normal max {ng[both].max():.2e} (>1e-3: {int((ng[both] > 1e-3).sum())} of {int(both.sum())})")


def test_trace_host_pipelined_equals_forward(fit3):
    """SphereTracer.trace_host (pinned host rays in, host RenderBuffer out, 3-stream chunk pipeline) returns exactly
    what forward() returns on the device: rays are independent, so chunking cannot change a single bit."""
    from nglod_b200.lib.tracer import SphereTracer
    net3, args3 = fit3_model(fit3, DEV)
    net3.lod = 2
    torch.manual_seed(5)
    o, d = O.look_at([-2.8, 2.8, -2.8], [0, 0, 0], 320, 180, fov=30.0)
    tr = SphereTracer(args3)
    ref = tr(net3, o.to(DEV), d.to(DEV))
    for chunks in (1, 3, 4):
        rb = tr.trace_host(net3, o.pin_memory(), d.pin_memory(), chunks=chunks)
        assert not rb.x.is_cuda
        assert torch.equal(rb.hit, ref.hit.cpu()) and torch.equal(rb.depth, ref.depth.cpu())
        assert torch.equal(rb.x, ref.x.cpu()) and torch.equal(rb.normal, ref.normal.cpu())
    rb = tr.trace_host(net3, o.pin_memory(), d.pin_memory(), fractions=(0.1, 0.3, 0.55, 0.8, 0.95), streams=3)
    assert torch.equal(rb.hit, ref.hit.cpu()) and torch.equal(rb.depth, ref.depth.cpu())
    with pytest.raises(RuntimeError):
        tr.trace_host(net3, o.to(DEV), d.to(DEV))


def test_trace_lookat_host_chunked_and_packed_equal_forward(fit3):
    """SphereTracer.trace_lookat_host (camera in, pinned host RenderBuffer out): the chunked pipeline, the 32-byte packed
    records and the 16-byte packed records + hit bytes (nglod_sphere_trace_packed writes PINNED HOST memory from the
    kernel) all return, bit for bit, what forward() returns for the rays `look_at` builds from the same window."""
    from nglod_b200.lib.tracer import SphereTracer
    from nglod_b200.lib.geoutils import _window
    from nglod_b200 import ops
    net3, args3 = fit3_model(fit3, DEV)
    net3.lod = 2
    W, H, cam, to = 160, 90, [-2.8, 2.8, -2.8], [0.0, 0.0, 0.0]
    torch.manual_seed(11)
    wx, wy = _window(W, H, "cpu")
    wx, wy = wx.pin_memory(), wy.pin_memory()
    tr = SphereTracer(args3)
    chunked = tr.trace_lookat_host(net3, cam, to, W, H, fov=30.0, window=(wx, wy), fields=("x", "depth", "hit", "normal"))
    # the same rays through forward(): regenerate them on the device from the same window
    from nglod_b200.lib.geoutils import camera_basis
    origin, view, right, up = camera_basis(cam, to)
    o, d = ops.generate_rays(origin, view, right, up, np.float32(np.tan(np.radians(15.0))),
                             False, wx.to(DEV), wy.to(DEV))
    ref = tr(net3, o, d)
    assert int(ref.hit.sum()) > 500
    for k in ("x", "depth", "hit", "normal"):
        assert torch.equal(getattr(chunked, k), getattr(ref, k).cpu()), k
    p32 = tr.trace_lookat_host(net3, cam, to, W, H, fov=30.0, window=(wx, wy), fields=("x", "depth", "hit", "normal"),
                               packed=True)
    for k in ("x", "depth", "hit", "normal"):
        assert not getattr(p32, k).is_cuda and torch.equal(getattr(p32, k), getattr(ref, k).cpu()), k
    out16 = {}
    for _ in range(2):                      # second call re-uses the pinned buffers
        p16 = tr.trace_lookat_host(net3, cam, to, W, H, fov=30.0, window=(wx, wy), out=out16, packed=True)
        assert p16.x is None and p16.depth.data_ptr() == out16["packed"].data_ptr() and out16["packed"].is_pinned()
        for k in ("depth", "hit", "normal"):
            assert torch.equal(getattr(p16, k), getattr(ref, k).cpu()), k
    # a capped grid (nglod_trace_opts_t.max_ctas: SMs left to other streams) traces the same frame
    capped = ops.sphere_trace(net3.net_view(), 2, o, d, max_ctas=5)
    for got, k in zip(capped, ("x", "depth", "hit", "normal")):
        assert torch.equal(got, getattr(ref, k)), k
    # device-resident packed buffers give the same records; argument checks of the C entry point
    view3 = net3.net_view()
    rec = ops.sphere_trace_packed(view3, 2, o, d, torch.empty(W * H, 8, device=DEV))
    assert torch.equal(rec.cpu(), out_packed32(ref))
    with pytest.raises(RuntimeError):
        ops.sphere_trace_packed(view3, 2, o, d, torch.empty(W * H, 8))                 # pageable host memory
    with pytest.raises(RuntimeError):
        ops.sphere_trace_packed(view3, 2, o, d, torch.empty(W * H, 4, device=DEV))     # 16-byte records need hit
    with pytest.raises(RuntimeError):
        ops.sphere_trace_packed(view3, 2, o, d, torch.empty(W * H * 8 + 1, device=DEV)[1:].view(W * H, 8))   # misaligned


def test_net_view_is_reused_and_never_stale():
    """OctreeSDF.net_view() hands the same borrowed view back while no parameter was written or moved, a fresh one after an
    optimiser step / mark_grids_dirty() / load_state_dict, and keeps the module deep-copyable and picklable."""
    import copy, io
    net, _ = rand5_model(DEV)
    x = torch.rand(4096, 3, device=DEV) * 2 - 1
    v0 = net.net_view()
    d0 = net.sdf(x, lod=4)
    assert net.net_view() is v0
    with torch.no_grad():
        net.features[4].fm.mul_(1.5)                       # version counter bumps
    v1 = net.net_view()
    assert v1 is not v0 and not torch.equal(net.sdf(x, lod=4), d0)
    net.features[4].fm.data.mul_(1.0 / 1.5)                # behind torch's back: needs mark_grids_dirty()
    net.mark_grids_dirty()
    assert net.net_view() is not v1
    assert (net.sdf(x, lod=4) - d0).abs().max().item() < 1e-6
    twin = copy.deepcopy(net)
    assert torch.equal(twin.sdf(x, lod=4), net.sdf(x, lod=4)) and twin.net_view() is not net.net_view()
    buf = io.BytesIO()
    torch.save(net, buf)
    buf.seek(0)
    back = torch.load(buf, weights_only=False)
    assert torch.equal(back.sdf(x, lod=4), net.sdf(x, lod=4))


def out_packed32(rb):
    n = rb.hit.shape[0]
    rec = torch.zeros(n, 8)
    rec[:, 0:1], rec[:, 1:4], rec[:, 5:8] = rb.depth.cpu(), rb.normal.cpu(), rb.x.cpu()
    rec.view(torch.int32)[:, 4] = rb.hit.cpu().to(torch.int32)
    return rec


# ------------------------------------------------------------------------------------------------ full-size properties
def test_full_size_properties_2pow20(rand5):
    """BASELINE configs[0] size (2^20 random points, lod 4), checked through size-independent properties:
    (1) the summed-grid and the per-LOD formulations agree, and so do the tensor-core and CUDA-core decoders;
    (2) queries are independent: evaluating two halves separately is bit-identical to one call;
    (3) backward: db1 == sum(grad_out), and -- trilinear weights and the restriction cascade are both partitions of
        unity -- the per-channel sum of dL/dfm_l is the same for EVERY l <= lod (it equals sum_q dL/dfeat_q)."""
    from nglod_b200 import ops
    net, _ = rand5_model(DEV)
    g = torch.Generator(device=DEV).manual_seed(1)
    n = 1 << 20
    x = torch.rand(n, 3, device=DEV, generator=g) * 2 - 1
    with torch.no_grad():
        net.math_mode, net.sum_lods = "tc", True
        a = net.sdf(x, lod=4)
        halves = torch.cat([net.sdf(x[: n // 2 + 17], lod=4), net.sdf(x[n // 2 + 17:], lod=4)])
        assert torch.equal(a, halves)
        net.sum_lods = False
        b = net.sdf(x, lod=4)
        net.math_mode, net.sum_lods = "fp32", True
        c = net.sdf(x, lod=4)
        print(f"2^20: summed vs per-LOD {float((a - b).abs().max()):.2e}; tc vs fp32 decoder {float((a - c).abs().max()):.2e}")
        assert (a - b).abs().max() < 2e-6 and (a - c).abs().max() < 5e-6
    net.math_mode, net.sum_lods = "tc", True
    go = torch.randn(n, 1, device=DEV, generator=g)
    for p in net.parameters():
        p.grad = None
    net.sdf(x, lod=4).backward(go)
    db1 = net.louts[4][2].bias.grad
    assert abs(float(db1) - float(go.double().sum())) < 1e-3 * float(go.abs().sum()) ** 0.5 + 1e-2
    sums = [net.features[l].fm.grad.double().sum(dim=(0, 2, 3, 4)) for l in range(5)]      # per channel
    scale = float(sums[4].abs().max())
    for l in range(4):
        assert float((sums[l] - sums[4]).abs().max()) < 2e-4 * scale + 1e-6, (l, float((sums[l] - sums[4]).abs().max()), scale)
    assert all(float(t.abs().max()) == 0.0 for t in net.summed_grad_scratch())


def test_realtime_loop_dense_and_sparse(fit3):
    """The headless display() loop (app/realtime.py): frame 0 uses the reference's default camera, so its hit mask must
    equal SphereTracer.forward on look_at rays with the same jittered window; the sparse mode must see the same object."""
    from nglod_b200.app import realtime
    from nglod_b200.lib.tracer import SphereTracer
    net3, args3 = fit3_model(fit3, DEV)
    torch.manual_seed(3)
    out = realtime.run(net3, 160, 90, frames=4, lod=2)
    assert len(out["ms"]) == 4 and out["rgb"].shape == (160 * 90, 3) and out["fps"] > 0
    torch.manual_seed(3)
    one = realtime.run(net3, 160, 90, frames=1, lod=2)
    torch.manual_seed(3)
    from nglod_b200.lib.geoutils import look_at
    o, d = look_at([-2.8, 2.8, -2.8], [0, 0, 0], 160, 90, mode="persp", fov=30.0, device=DEV)
    net3.lod = 2
    ref = SphereTracer(args3)(net3, o, d)
    assert int((one["hit"] != ref.hit).sum()) <= 2          # eye = radius*(cos, sin) of 225 deg vs the literal (-2.8, -2.8)
    assert int(ref.hit.sum()) > 300
    rgb = one["rgb"]
    assert float(rgb.min()) >= 0.0 and float(rgb.max()) <= 1.0 + 1e-6
    assert bool((rgb[~one["hit"]] == 1.0).all())
    torch.manual_seed(3)
    sp = realtime.run(net3, 160, 90, frames=2, lod=2, spc_level=4)
    assert "sparse" in sp["mode"]
    inter = int((sp["hit"] & out["hit"]).sum()) if sp["hit"].shape == out["hit"].shape else 0
    assert int(sp["hit"].sum()) > 300


# ------------------------------------------------------------------------------------------------ model variants
@pytest.mark.parametrize("flags,kw", [
    (["--num-lods", "3", "--pos-invariant"], dict(pos_invariant=True)),
    (["--num-lods", "4", "--joint-decoder"], {}),
    (["--num-lods", "3", "--base-lod", "1"], {}),
    (["--num-lods", "2", "--base-lod", "3"], {}),
    (["--num-lods", "7", "--base-lod", "0"], {}),         # R = 1 .. 64: more grids than one set-up batch (TC_PACK_LODS = 5)
    # models smaller than the kernels' 32 / 128 (the reference README's `--feature-dim 16`): run zero-padded, exactly
    (["--num-lods", "3", "--feature-dim", "16"], {}),
    (["--num-lods", "3", "--feature-dim", "16", "--hidden-dim", "64", "--joint-decoder"], {}),
    (["--num-lods", "2", "--feature-dim", "8", "--hidden-dim", "32", "--pos-invariant"], dict(pos_invariant=True)),
])
def test_model_variants_forward_backward_trace(flags, kw):
    """The reference's other OctreeSDF shapes (--pos-invariant, --joint-decoder, --base-lod, up to 7 LODs) through every
    path: both decoders x (summed grid | per-LOD gather), autograd backward on both scatter paths, a short trace."""
    from nglod_b200.lib.tracer import SphereTracer
    args = make_args(flags)
    torch.manual_seed(3)
    net = OctreeSDF_cls(args).to(DEV)
    for f in net.features:                       # larger features than the 0.01 init so every LOD matters
        f.fm.data.mul_(20.0)
    net.mark_grids_dirty()
    onet = O.OracleNet(net.state_dict(), requires_grad=True, **kw)
    g = torch.Generator().manual_seed(5)
    x = torch.cat([torch.rand(3001, 3, generator=g) * 2.3 - 1.15, torch.tensor([[1.0, 1.0, 1.0], [-1.0, -1.0, 1.0]])])
    top = net.num_lods - 1
    with torch.no_grad():
        for lod in sorted({0, top // 2, top}):
            ref = onet.sdf(x, lod=lod)
            for mode in ("tc", "fp32"):
                for summ in (True, False):
                    net.math_mode, net.sum_lods = mode, summ
                    err = (net.sdf(x.to(DEV), lod=lod).cpu() - ref).abs().max().item()
                    assert err < 1e-5, (flags, lod, mode, summ, err)
    # fp16 x-pair lines of the summed grid (grids as small as R = 1): equals the oracle on the model's summed_state_dict
    net.math_mode, net.sum_lods, net.grid_storage = "tc", True, "fp16"
    with torch.no_grad():
        for lod in sorted({0, top}):
            oq = O.OracleNet(net.summed_state_dict(lod), **kw)
            assert (net.sdf(x.to(DEV), lod=lod).cpu() - oq.sdf(x, lod=lod)).abs().max() < 1e-5, (flags, lod, "fp16")
    net.grid_storage = "fp32"
    net.math_mode = "tc"
    gt = torch.rand(x.shape[0], 1, generator=g) - 0.5
    loss_ref = ((onet.sdf(x, lod=top) - gt) ** 2).mean()
    loss_ref.backward()
    for summ in (True, False):
        net.sum_lods = summ
        for p in net.parameters():
            p.grad = None
        loss = ((net.sdf(x.to(DEV), lod=top) - gt.to(DEV)) ** 2).mean()
        loss.backward()
        assert abs(loss.item() - loss_ref.item()) < 1e-5 * max(1.0, abs(loss_ref.item()))
        for i in range(net.num_lods):
            a, b = net.features[i].fm.grad.cpu(), onet.fm[i].grad
            assert (a - b).abs().max() / b.abs().max() < 3e-4, (flags, summ, i)
        dec = net.decoder_params(top)
        for a, b in zip(dec, onet.decoder(top)):
            assert (a.grad.cpu() - b.grad).abs().max() / b.grad.abs().max() < 3e-4, (flags, summ)
    net.sum_lods, net.lod, onet.lod = True, top, top
    torch.manual_seed(4)
    o, d = O.look_at([-2.8, 2.8, -2.8], [0, 0, 0], 64, 36, fov=30.0)
    rb = SphereTracer(args, num_steps=24)(net, o.to(DEV), d.to(DEV))
    with torch.no_grad():
        ref = O.sphere_trace(onet, o, d, num_steps=24)
    assert int((rb.hit.cpu() != ref["hit"]).sum()) <= 2


def test_mesh2sdf_scratch_pool_can_be_released():
    """nglod_release_scratch: the library's one piece of process-wide state (the per-device mesh2sdf scratch pool) hands
    its memory back on request, and the next call simply re-grows it with identical results."""
    from nglod_b200 import ops
    from nglod_b200.lib.torchgp import icosphere, point_sample, normalize
    V, F = normalize(*[t.to(DEV) for t in icosphere(4)])
    tri = V[F].contiguous()
    torch.manual_seed(1)
    pts = point_sample(V, F, ["rand", "near", "trace"], 20000)
    a = ops.mesh2sdf_gpu(pts, tri)[0].clone()
    torch.cuda.synchronize()
    ops.release_scratch()
    b = ops.mesh2sdf_gpu(pts, tri)[0]
    assert torch.equal(a, b)
    ops.release_scratch()
