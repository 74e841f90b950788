"""CPU: host-side logic and the C-ABI surface (no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from helpers import make_args, rand5_model

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_cabi_library_exports_every_declared_symbol():
    from nglod_b200 import _lib
    header = open(os.path.join(ROOT, "include", "nglod_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|const char\*)\s+(nglod_\w+)\s*\(", header, flags=re.M))
    assert len(declared) >= 12
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.nglod_abi_version() == _lib.EXPECTED_ABI
    assert b"sm_100a" in lib.nglod_build_info()
    # struct layouts agree with the header (sizes computed from the declarations)
    assert ctypes.sizeof(_lib.NetStruct) == 6 * 4 + 8 * 4 + 7 * 8 * 8
    assert ctypes.sizeof(_lib.NetGradStruct) == 6 * 8 * 8 + 8 + 8
    assert ctypes.sizeof(_lib.TraceOpts) == 8 + 4 * 8 + 8          # + max_ctas, reserved_ (ABI 12)
    # argument validation happens before any CUDA call, so it is testable without a device
    assert lib.nglod_aabb(None, None, -1, None, None, None, None) == _lib.EINVAL
    assert lib.nglod_sdf_forward(None, 0, None, 0, None, None) == _lib.EINVAL
    assert lib.nglod_aabb(None, None, 0, None, None, None, None) == 0


def test_no_cpu_fallback():
    from nglod_b200 import ops
    net, _ = rand5_model("cpu")
    with pytest.raises(RuntimeError, match="CUDA"):
        net.sdf(torch.zeros(3, 3), lod=0)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.aabb(torch.zeros(3, 3), torch.ones(3, 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.mesh2sdf_gpu(torch.zeros(3, 3), torch.zeros(2, 3, 3))


def test_product_never_imports_oracle():
    for base, _, files in os.walk(os.path.join(ROOT, "nglod_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(base, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), os.path.join(base, f)


def test_options_match_reference_defaults():
    from nglod_b200.lib.options import parse_options, argparse_to_str
    p = parse_options(return_parser=True)
    a = p.parse_args([])
    assert (a.net, a.feature_dim, a.num_lods, a.base_lod, a.hidden_dim) == ("OverfitSDF", 32, 1, 2, 128)
    assert a.sample_mode == ["rand", "near", "near", "trace", "trace"] and a.num_samples == 100000
    assert (a.num_steps, a.step_size, a.min_dis, a.camera_clamp) == (256, 1.0, 0.0003, [-5, 10])
    assert a.camera_origin == [-2.8, 2.8, -2.8] and a.camera_fov == 30 and a.render_res == [512, 512]
    assert a.grad_method == "finitediff" and a.tracer == "SphereTracer" and a.lod is None and a.lr == 0.001
    a = p.parse_args("--net OctreeSDF --num-lods 5 --lod 4 --render-res 1280 720 --shadow --ao".split())
    assert a.lod == 4 and a.render_res == [1280, 720] and a.shadow and a.ao
    grp = p.add_argument_group("app")                 # apps extend the parser (sdf_renderer.py:59-82)
    grp.add_argument("--img-dir", type=str, default="x")
    args, txt = argparse_to_str(p, [])
    assert txt.startswith("```") and "'renderer'" in txt and args.img_dir == "x"


def test_setparam_and_tracer_ctor():
    from nglod_b200.lib.utils import setparam
    from nglod_b200.lib.tracer import SphereTracer
    a = make_args()
    assert setparam(a, None, "num_steps") == 256 and setparam(a, 12, "num_steps") == 12
    assert setparam(None, None, "num_steps") is None
    t = SphereTracer(a, num_steps=12)
    assert t.num_steps == 12 and t.min_dis == 0.0003 and t.camera_clamp == [-5, 10] and t.grad_method == "finitediff"
    t = SphereTracer(camera_clamp=[0, 5], step_size=0.5, grad_method="finitediff", num_steps=8, min_dis=1e-3)
    assert t.step_size == 0.5 and t.inv_num_steps == 1 / 8


def test_renderbuffer_semantics():
    from nglod_b200.lib.tracer import RenderBuffer
    a = RenderBuffer(x=torch.zeros(4, 3), hit=torch.ones(4, dtype=torch.bool))
    b = RenderBuffer(x=torch.ones(2, 3), depth=torch.ones(2, 1))
    c = a + b
    assert c.x.shape == (6, 3) and c.hit.shape == (4,) and c.depth.shape == (2, 1) and c.normal is None
    e = RenderBuffer()
    e += a
    assert e.x.shape == (4, 3)
    assert len(list(a)) == 12
    r = RenderBuffer(rgb=torch.rand(6, 3), hit=torch.ones(6, 1)).reshape(3, 2, -1).transpose()
    assert r.rgb.shape == (2, 3, 3) and r.hit.shape == (2, 3, 1)
    d = r.exrdict()
    assert "default" in d and "rgb" not in d and "x" not in d
    img = RenderBuffer(hit=torch.ones(2, 2, 1), normal=torch.zeros(2, 2, 3), rgb=torch.ones(2, 2, 3) * 0.5,
                       relative_depth=torch.ones(2, 2, 1) * 0.2).image()
    assert img.hit.shape == (2, 2, 3) and float(img.normal[0, 0, 0]) == 127.5 and float(img.depth[0, 0, 0]) == 51.0
    m = RenderBuffer.mean(RenderBuffer(rgb=torch.zeros(2, 3)), RenderBuffer(rgb=torch.ones(2, 3)))
    assert torch.allclose(m.rgb, torch.full((2, 3), 0.5))


def test_reference_checkpoint_roundtrip():
    """A reference-layout state_dict (contiguous NCDHW grids) loads; ours saves back with the same keys/shapes;
    the grids stay channels-last physically."""
    net, _ = rand5_model()
    sd = {k: v.clone().contiguous() for k, v in net.state_dict().items()}
    args = make_args(["--num-lods", "5"])
    from nglod_b200.lib.models import OctreeSDF
    net2 = OctreeSDF(args)
    net2.load_state_dict(sd)
    for i in range(5):
        fm = net2.features[i].fm
        assert fm.is_contiguous(memory_format=torch.channels_last_3d)
        assert torch.equal(fm, sd[f"features.{i}.fm"])
        assert torch.equal(fm.permute(0, 2, 3, 4, 1).contiguous()[0, 1, 2, 3], sd[f"features.{i}.fm"][0, :, 1, 2, 3])
    assert list(net2.state_dict().keys()) == list(sd.keys())


def test_octree_sdf_ctor_variants():
    from nglod_b200.lib.models import OctreeSDF
    n = OctreeSDF(make_args(["--num-lods", "3", "--joint-decoder"]))
    assert len(n.louts) == 1 and n.num_decoder == 1
    n = OctreeSDF(make_args(["--num-lods", "2", "--pos-invariant"]))
    assert n.louts[0][0].weight.shape == (128, 32)
    n = OctreeSDF(make_args(["--num-lods", "2", "--base-lod", "3"]))
    assert n.features[0].fm.shape[-1] == 9
    with pytest.raises(NotImplementedError):
        OctreeSDF(make_args(["--num-lods", "2", "--pos-enc"]))
    n = OctreeSDF(make_args(["--num-lods", "2", "--interpolate", "0.5"]))
    n.lod = 0
    with pytest.raises(NotImplementedError):
        n.sdf(torch.zeros(1, 3))
    n.freeze()
    assert not any(p.requires_grad for p in n.parameters())


def test_look_at_matches_oracle_with_same_seed():
    from nglod_b200.lib.geoutils import look_at
    from oracle import nglod_oracle as O
    for mode in ("persp", "ortho"):
        torch.manual_seed(9)
        o1, d1 = look_at([-2.8, 2.8, -2.8], [0, 0, 0], 40, 30, mode=mode, fov=30.0, device="cpu")
        torch.manual_seed(9)
        o2, d2 = O.look_at([-2.8, 2.8, -2.8], [0, 0, 0], 40, 30, mode=mode, fov=30.0)
        assert torch.equal(o1, o2) and torch.equal(d1, d2)
    assert o1.shape == (1200, 3)
    with pytest.raises(ValueError):
        look_at([0, 0, 1], [0, 0, 0], 4, 4, mode="fisheye", device="cpu")


def test_matcap_and_blur_match_scipy():
    from scipy.interpolate import RegularGridInterpolator
    from scipy.ndimage import gaussian_filter
    from nglod_b200.lib.geoutils import MatcapSampler, gaussian_blur2d, spherical_envmap
    rng = np.random.RandomState(0)
    tex = rng.rand(17, 23, 3).astype(np.float32) * 255
    uv = rng.rand(500, 2)
    uv[:4] = [[0, 0], [1, 1], [0, 1], [1, 0]]
    ref = RegularGridInterpolator((np.linspace(0, 1, 17), np.linspace(0, 1, 23)), tex)(uv)
    got = MatcapSampler(torch.from_numpy(tex))(torch.from_numpy(uv).float()).numpy()
    assert np.abs(got - ref).max() < 1e-2            # fp32 vs fp64 on values up to 255
    img = rng.rand(31, 18).astype(np.float32)
    assert np.abs(gaussian_blur2d(torch.from_numpy(img), 2.0).numpy() - gaussian_filter(img, sigma=2)).max() < 1e-6
    small = rng.rand(5, 40).astype(np.float32)       # image narrower than the 8-pixel kernel radius
    assert np.abs(gaussian_blur2d(torch.from_numpy(small), 2.0).numpy() - gaussian_filter(small, sigma=2)).max() < 1e-6
    n = torch.nn.functional.normalize(torch.randn(50, 3), dim=1)
    v = torch.nn.functional.normalize(torch.randn(50, 3), dim=1)
    uvm = spherical_envmap(v, n)
    assert uvm.shape == (50, 2) and float(uvm.min()) >= 0 and float(uvm.max()) <= 1


def test_samplers_and_meshes():
    from nglod_b200.lib.torchgp import (icosphere, torus, point_sample, sample_surface, normalize,
                                        area_weighted_distribution, load_obj, write_obj)
    V, F = icosphere(2)
    assert V.shape == (162, 3) and F.shape == (320, 3)
    assert torch.allclose(V.norm(dim=1), torch.ones(162), atol=1e-6)
    tri = V[F]
    nrm = torch.linalg.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    assert ((nrm * tri.mean(dim=1)).sum(-1) > 0).all()            # outward winding
    Vt, Ft = torus(nu=16, nv=8)
    assert Vt.shape == (128, 3) and Ft.shape == (256, 3)
    # closed manifold: every edge is shared by exactly two triangles
    e = torch.cat([Ft[:, [0, 1]], Ft[:, [1, 2]], Ft[:, [2, 0]]]).sort(dim=1)[0]
    _, counts = torch.unique(e, dim=0, return_counts=True)
    assert (counts == 2).all()
    Vn, _ = normalize(Vt * 3 + 1.5, Ft)
    assert abs(float(Vn.norm(dim=1).max()) - 1.0) < 1e-6
    torch.manual_seed(0)
    pts = point_sample(V, F, ["rand", "near", "trace"], 1000)
    assert pts.shape == (3000, 3)
    assert (pts[:1000].abs() <= 1).all()
    assert (pts[2000:].norm(dim=1) <= 1 + 1e-6).all() and (pts[2000:].norm(dim=1) > 0.93).all()
    s, n = sample_surface(V, F, 10)
    assert s.shape == (10, 3) and n.shape == (10, 3)
    dist = area_weighted_distribution(V, F)
    assert abs(float(dist.probs.sum()) - 1) < 1e-5
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "m.obj")
        write_obj(path, V, F)
        V2, F2 = load_obj(path)
        assert torch.allclose(V2, V, atol=1e-6) and torch.equal(F2, F)
        with open(path, "a") as fh:
            fh.write("f 1/1/1 2/2/2 3/3/3 4/4/4\n")              # a quad with v/vt/vn indices
        _, F3 = load_obj(path)
        assert F3.shape[0] == F.shape[0] + 2


def test_host_samplers_reproduce_the_reference_under_equal_seeds():
    """tests/golden/samplers.npz was produced by the reference's own lib/torchgp functions (imported in place) under fixed
    torch seeds; the CPU path of nglod_b200.lib.torchgp is the same recipe and must give the same bits.  (The sampler
    kernel draws from Philox instead and is compared with this recipe distributionally in the GPU tests.)"""
    import os
    import numpy as np
    from nglod_b200.lib.torchgp import (normalize, per_face_normals, point_sample, sample_surface, sample_near_surface,
                                        sample_spc)
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "samplers.npz"))
    V, F = torch.from_numpy(g["V"]), torch.from_numpy(g["F"])
    Vn, Fn = normalize(V.clone(), F.clone())
    assert np.array_equal(Vn.numpy(), g["Vn"])
    assert np.array_equal(per_face_normals(Vn, Fn).numpy(), g["face_normals"])
    torch.manual_seed(7)
    assert np.array_equal(point_sample(Vn, Fn, ["rand", "near", "trace", "near"], 64).numpy(), g["point_sample"])
    torch.manual_seed(8)
    sp, sn = sample_surface(Vn, Fn, 50)
    assert np.array_equal(sp.numpy(), g["surface_pts"]) and np.array_equal(sn.numpy(), g["surface_nrm"])
    torch.manual_seed(9)
    assert np.array_equal(sample_near_surface(Vn, Fn, 50, variance=0.02).numpy(), g["near"])
    torch.manual_seed(10)
    assert np.array_equal(sample_spc(torch.from_numpy(g["spc_corners"]), 3, 6).numpy(), g["spc_samples"])


def test_shard_helpers():
    from nglod_b200.dist import shard_range, interleaved_strips
    for n, w, al in ((921600, 8, 720), (10, 4, 1), (7, 8, 1), (1000, 3, 16)):
        cover = []
        for r in range(w):
            s, e = shard_range(n, r, w, al)
            assert 0 <= s <= e <= n and (s % al == 0)
            cover += list(range(s, e))
        assert cover == list(range(n))
    cols = sorted(c for r in range(4) for c0, c1 in interleaved_strips(1280, r, 4) for c in range(c0, c1))
    assert cols == list(range(1280))
    assert len(interleaved_strips(1280, 0, 4, strips_per_rank=4)) == 4


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) must run without a GPU and print exactly
    one JSON line on stdout carrying the contract's keys."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, NGLOD_FIT_STEPS="2", NGLOD_REF_BUDGET_S="0.05", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "rays/s" and j["value"] > 0
    # "reference" = the unmodified reference classes (from /root/reference or their staged copy under oracle/_ref) on the
    # host cores; "port" = the oracle restatement, only when neither is present
    assert j["cpu_baseline"]["kind"] in ("reference", "port") and j["e2e"]["h2d_bytes_per_step"] == 0
    assert j["config"]["workload"].startswith("SphereTracer.forward 1280x720") and "details" in j


def test_bench_reference_arm_other_ranks_exit_without_work():
    """Under torchrun (N > 1) rank 0 alone runs and prints the reference arm; every other rank exits 0 with nothing on
    stdout and without joining a process group."""
    import subprocess
    import sys
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29999",
               CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip() == ""


def test_camera_basis_equals_torch_recipe():
    """geoutils.camera_basis (nglod_camera_basis, host-only C) == look_at's torch recipe on the host bit for bit
    (reference geoutils.py:180-188), on random poses and the degenerate straight-down view."""
    import torch.nn.functional as F
    from nglod_b200.lib.geoutils import camera_basis

    def recipe(f, t):
        origin = torch.tensor(list(f), dtype=torch.float32)
        view = F.normalize(torch.tensor(list(t), dtype=torch.float32) - origin, dim=0)
        right = F.normalize(torch.linalg.cross(view, torch.tensor([0.0, 1.0, 0.0])), dim=0)
        up = F.normalize(torch.linalg.cross(right, view), dim=0)
        return origin.tolist(), view.tolist(), right.tolist(), up.tolist()

    rng = np.random.default_rng(3)
    poses = [((rng.standard_normal(3) * 3).tolist(), (rng.standard_normal(3) * 0.3).tolist()) for _ in range(3000)]
    poses += [([-2.8, 2.8, -2.8], [0, 0, 0]), ([0.0, 3.0, 0.0], [0.0, 0.0, 0.0]), ([1.0, 1.0, 1.0], [1.0, 1.0, 1.0])]
    for f, t in poses:
        got, want = camera_basis(f, t), recipe(f, t)
        for g, w in zip(got, want):
            assert np.array_equal(np.array(g, dtype=np.float32), np.array(w, dtype=np.float32), equal_nan=True), (f, t, got, want)


def test_header_is_plain_c_and_a_c_client_links(tmp_path):
    """include/nglod_b200.h compiles as C99 (what a cgo / JNI / ctypes-less client sees), a C program links against
    libnglod_b200.so and calls the entry points that do no device work, and the struct sizes the C compiler sees are the
    ones the ctypes mirror in nglod_b200/_lib.py declares (a layout drift would make the kernels read wrong pointers)."""
    import subprocess
    from nglod_b200 import _lib
    from nglod_b200.build import build_library
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = build_library()
    exe = tmp_path / "client"
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(root, "include"),
           os.path.join(root, "tests", "c_client", "client.c"), "-o", str(exe), so, "-lm", f"-Wl,-rpath,{os.path.dirname(so)}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    m = re.search(r"ok abi (\d+) sizeof\(nglod_net_t\) (\d+) sizeof\(nglod_net_grad_t\) (\d+) sizeof\(nglod_trace_opts_t\) (\d+) sizeof\(nglod_sparse_net_t\) (\d+)", r.stdout)
    assert m, r.stdout
    assert int(m.group(1)) == _lib.EXPECTED_ABI
    assert int(m.group(2)) == ctypes.sizeof(_lib.NetStruct)
    assert int(m.group(3)) == ctypes.sizeof(_lib.NetGradStruct)
    assert int(m.group(4)) == ctypes.sizeof(_lib.TraceOpts)
    assert int(m.group(5)) == ctypes.sizeof(_lib.SparseNetStruct)
    offs = [int(v) for v in re.search(r"offsets ([\d ]+)", r.stdout).group(1).split()]
    assert offs == [_lib.NetStruct.grid_res.offset, _lib.NetStruct.grids.offset, _lib.NetStruct.w0.offset,
                    _lib.NetStruct.summed_fp16.offset, _lib.NetGradStruct.summed.offset,
                    _lib.NetGradStruct.scatter_scratch_floats.offset, _lib.TraceOpts.step_size.offset,
                    _lib.TraceOpts.normal_h.offset]


def test_committed_bench_record_has_the_contract_keys():
    """The last bench line committed under profiles/ carries every key of the measurement contract (driver-parsed fields, the
    roofline / cpu_baseline / e2e objects, clocks, launches), with the relations between them that the contract states."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = json.loads(open(os.path.join(root, "profiles", "bench_r2_n1.json")).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "l2" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] >= d["steps"]
    rays = d["config"]["rays_per_gpu_step"]
    assert abs(d["value"] - rays / (d["ms_per_step"] / 1e3)) / d["value"] < 1e-6          # value = rays / device time per step
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.0 < r["frac"] < 1.2
    c = d["cpu_baseline"]
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in c, k
    assert c["kind"] in ("reference", "port") and c["unit"] == d["unit"]
    e = d["e2e"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in e, k
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] >= rays * 17 and e["value"] < d["value"]
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(d["clocks"])
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_window_cache_draws_and_rounds_like_normalized_grid():
    """geoutils._window (cached un-jittered coordinates + one torch.rand per axis) == the window vectors of normalized_grid
    (reference geoutils.py:140-154) under the same seed, bit for bit, on the first (cache miss) and on later (cache hit) calls."""
    from nglod_b200.lib.geoutils import _window, normalized_grid
    for w, h in ((160, 90), (33, 57), (160, 90)):
        torch.manual_seed(5)
        grid = normalized_grid(w, h, device="cpu")
        torch.manual_seed(5)
        wx, wy = _window(w, h, "cpu")
        assert torch.equal(wx, grid[:, 0, 0]) and torch.equal(wy, grid[0, :, 1])
