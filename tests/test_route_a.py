"""INTEGRATION.md route A: the UNMODIFIED reference Python (sdf-net/lib, staged under oracle/_ref by oracle/build_ref.py)
running over nglod_b200/shims -- its `from sol_nglod import aabb` and `import mesh2sdf` bind to the sm_100a kernels --
next to the package's own classes on identical weights, rays and points.  Runs in a subprocess per binding because the
reference caches its extension imports at module load."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
pytestmark = pytest.mark.gpu

SCRIPT = r'''
import json, sys, numpy as np, torch
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
from oracle import ref_python
ref = ref_python.import_reference(%(ext)r)
from helpers import fit3_model, make_args
from nglod_b200.lib.tracer import SphereTracer as OurTracer
from nglod_b200.lib.torchgp import torus
from nglod_b200 import ops
fit3 = dict(np.load(%(root)r + "/tests/golden/fit3.npz"))
dev = "cuda"
# ---- the reference's OctreeSDF + SphereTracer (pure torch on the GPU + aabb from the bound extension)
rargs = ref.parse_options(return_parser=True).parse_args(["--net", "OctreeSDF", "--num-lods", "3", "--feature-dim", "32"])
rnet = ref.OctreeSDF(rargs)
rnet.load_state_dict({k[3:]: torch.from_numpy(v) for k, v in fit3.items() if k.startswith("sd.")})
rnet = rnet.to(dev).eval(); rnet.lod = 2
ours, oargs = fit3_model(fit3, dev); ours.lod = 2
o = torch.from_numpy(fit3["t1_ray_o"]).to(dev); d = torch.from_numpy(fit3["t1_ray_d"]).to(dev)
with torch.no_grad():
    rrb = ref.SphereTracer(rargs)(rnet, o, d)
    orb = OurTracer(oargs)(ours, o, d)
    x = torch.rand(20000, 3, device=dev) * 2 - 1
    sdf_err = float((rnet.sdf(x, lod=2) - ours.sdf(x, lod=2)).abs().max())
conv = (rnet(rrb.x).abs() < 0.0003)[:, 0] & rrb.hit & orb.hit
res = {"hits_ref": int(rrb.hit.sum()), "hit_mismatch": int((rrb.hit != orb.hit).sum()), "sdf_err": sdf_err,
       "depth_err_conv": float((rrb.depth - orb.depth).abs()[:, 0][conv].max()),
       "normal_bad_frac": float(((rrb.normal - orb.normal).abs().max(dim=1)[0][conv] > 1e-3).float().mean()),
       "golden_hit_mismatch": int((rrb.hit.cpu().numpy() != fit3["t1_hit"]).sum())}
# ---- the reference's MeshDataset.resample (point_sample on the host + compute_sdf -> mesh2sdf.mesh2sdf_gpu)
V, F = torus(0.6, 0.25, 48, 24)
ds = ref.MeshDataset.__new__(ref.MeshDataset)
ds.args, ds.sample_mode, ds.get_normals, ds.num_samples, ds.sample_tex = None, ["rand", "near", "trace"], False, 4000, False
ds.V, ds.F = ref.torchgp.normalize(V, F)
torch.manual_seed(3)
ds.resample()
tri = ds.V[ds.F].to(dev).contiguous()
mine = ops.mesh2sdf_gpu(ds.pts.to(dev).contiguous(), tri)[0].cpu()
res["m2s_n"] = int(ds.pts.shape[0]); res["m2s_err"] = float((mine - ds.d[:, 0]).abs().max())
res["m2s_sign_flips"] = int(((mine < 0) != (ds.d[:, 0] < 0)).sum())
res["m2s_trace_on_surface"] = float(ds.d[8000:].abs().max())
print("RESULT " + json.dumps(res))
'''


def _run(ext):
    src = os.path.join(ROOT, "oracle", "_ref", "sdf-net", "lib")
    assert os.path.isdir(src) or os.path.isdir("/root/reference/sdf-net/lib"), \
        "reference Python is not staged: run `python oracle/build_ref.py` where /root/reference exists"
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT, "ext": ext}], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[7:])


def test_unmodified_reference_python_over_our_kernels():
    res = _run("ours")
    print("route A (reference Python over nglod_b200 shims):", res)
    assert res["hits_ref"] > 300
    # the reference's torch path on a GPU contracts x = o + d t into an fma (CUDA addcmul), the golden vectors are CPU:
    # threshold-straddling rays may flip, nothing else
    assert res["hit_mismatch"] <= 2 and res["golden_hit_mismatch"] <= 2
    assert res["sdf_err"] < 3e-6
    assert res["depth_err_conv"] < 2e-4 and res["normal_bad_frac"] < 0.01
    assert res["m2s_n"] == 12000 and res["m2s_err"] == 0.0 and res["m2s_sign_flips"] == 0      # same kernel on both sides
    assert res["m2s_trace_on_surface"] < 1e-4


def test_whole_reference_on_this_gpu_agrees_with_ours():
    """The same script with the reference's OWN compiled extensions bound: the complete reference on the B200."""
    res = _run("reference")
    print("reference (own kernels) on this GPU vs ours:", res)
    assert res["hit_mismatch"] <= 2 and res["golden_hit_mismatch"] <= 2
    assert res["sdf_err"] < 3e-6 and res["depth_err_conv"] < 2e-4
    assert res["m2s_err"] < 2e-6 and res["m2s_sign_flips"] == 0
