"""Profiling target for the host-to-host frame: SphereTracer.trace_lookat_host(packed=True) a few times (for ncu; never a
bench number).  The tracer launch writes its 16-byte records + hit bytes into PINNED HOST memory.
  ncu --set full --clock-control none --import-source on -k regex:sphere_trace_kernel -s 2 -c 1 -o gpurun_out/r2_e2e \
      python profiles/prof_e2e_target.py"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import bench  # noqa: E402
from nglod_b200.lib.tracer import SphereTracer  # noqa: E402
from nglod_b200.lib.geoutils import _window  # noqa: E402

bench.FIT_STEPS = int(os.environ.get("PROF_FIT_STEPS", "150"))
dev = torch.device("cuda", 0)
net, args = bench.build_and_fit(dev, lambda m: print(m, file=sys.stderr))
tracer = SphereTracer(args)
torch.manual_seed(1000)
wx, wy = _window(bench.W, bench.H, "cpu")
wx, wy = wx.pin_memory(), wy.pin_memory()
out = {}
for _ in range(4):
    rb = tracer.trace_lookat_host(net, bench.camera_from(0.0), bench.CAM_TO, bench.W, bench.H, fov=bench.FOV, window=(wx, wy),
                                  out=out, packed=True)
print("done, hits", int(rb.hit.sum()), file=sys.stderr)
