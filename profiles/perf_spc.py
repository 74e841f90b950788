"""BASELINE config 4 frame (1920x1080, level-7 octree of the torus, NeuralSPC with 6 LODs fitted for a few steps): per LOD
the traversal alone, the in-voxel tracer alone (on a resident nugget list) and the whole `trace()` call; a checksum of
the results so that variants can be compared."""
import sys, torch, numpy as np, ctypes
sys.path.insert(0, '/root/repo')
import bench
from nglod_b200 import _lib
from nglod_b200.lib import spc as S
from nglod_b200.lib.geoutils import look_at
from nglod_b200.lib.torchgp import torus, normalize
dev = torch.device('cuda', 0)
V, F = normalize(*[t.to(dev) for t in torus(0.6, 0.25, 128, 64)])
torch.manual_seed(77)
spc = S.SPC(S.mesh_to_octree(V, F, 7, num_samples=1 << 22))
torch.manual_seed(5)
ro, rd = look_at(bench.CAM_FROM, bench.CAM_TO, 1920, 1080, mode="persp", fov=bench.FOV, device=dev)
nspc = S.NeuralSPC(spc, num_lods=6, base_lod=2)
opt = torch.optim.Adam(nspc.parameters(), lr=1e-3)
gs = torch.Generator(device=dev).manual_seed(11)
lp7 = spc.level_points(7)[:, :3].float()
for it in range(120):
    pv = torch.randint(0, lp7.shape[0], (65536,), device=dev, generator=gs)
    xs7 = ((lp7[pv] + torch.rand(pv.shape[0], 3, device=dev, generator=gs)) / 128 * 2 - 1).contiguous()
    gt7 = (torch.sqrt((torch.sqrt(xs7[:, 0] ** 2 + xs7[:, 2] ** 2) - 0.6) ** 2 + xs7[:, 1] ** 2) - 0.25).unsqueeze(1)
    opt.zero_grad(set_to_none=False); nspc.loss_backward(xs7, gt7); opt.step()
nspc.eval()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, it=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(it):
        flush.zero_(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
with torch.no_grad():
    for lod in (1, 2, 3, 4, 5):
        x_, t_, hit_, n_, p_ = nspc.trace(ro, rd, lod)
        chk = (float(t_[hit_].double().sum()), int(hit_.sum()), int(p_.long().clamp(min=0).sum()), float(n_.double().abs().sum()))
        t_all = timed(lambda: nspc.trace(ro, rd, lod))
        t_trav = timed(lambda: spc.raytrace(ro, rd, lod + 2, return_offsets=True))
        print(f"lod {lod} (level {lod + 2}): trace() {t_all:.3f} ms | traversal {t_trav:.3f} ms | in-voxel tracer ~{t_all - t_trav:.3f} ms | "
              f"hits {chk[1]} depth-sum {chk[0]:.6f} pidx-sum {chk[2]} |n|-sum {chk[3]:.4f}", flush=True)
