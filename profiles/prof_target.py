"""Profiling target: a short in-run fit, then each hot kernel a few times (for ncu; never a bench number).
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      -k regex:"sphere_trace|sdf_|aabb|mesh2sdf|adam" python profiles/prof_target.py
  ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 2 -o gpurun_out/<name> python profiles/prof_target.py
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import bench  # noqa: E402
from nglod_b200 import ops  # noqa: E402
from nglod_b200.lib.tracer import SphereTracer  # noqa: E402

bench.FIT_STEPS = int(os.environ.get("PROF_FIT_STEPS", "150"))
dev = torch.device("cuda", 0)
net, args = bench.build_and_fit(dev, lambda m: print(m, file=sys.stderr))
net.grid_storage = os.environ.get("PROF_STORAGE", "fp32")      # "fp16": the x-pair-line inference path
view = net.net_view()
ray_o, ray_d = bench.make_rays(dev)
tracer = SphereTracer(args)
g = torch.Generator(device=dev).manual_seed(1)
xq = torch.rand(bench.SDF_N, 3, device=dev, generator=g) * 2 - 1
gq = torch.rand(bench.SDF_N, device=dev, generator=g)
grid_grads = [torch.zeros_like(f.fm.data, memory_format=torch.preserve_format) for f in net.features]
dec_grad = tuple(torch.zeros_like(p) for p in net.decoder_params(bench.LOD))
for _ in range(3):
    ops.aabb(ray_o, ray_d)
    ops.sdf_forward(view, bench.LOD, xq)
    ops.sdf_backward(view, bench.LOD, xq, gq, grid_grads, dec_grad, summed_scratch=net.summed_grad_scratch())
    tracer(net, ray_o, ray_d)
torch.cuda.synchronize()
print("done", file=sys.stderr)
