"""Build with -DNGLOD_WS_TIMING: every CTA of the warp-specialised forward prints when it finished (globaltimer).  How far
apart do the SMs finish a statically partitioned 2^20 / 2^23-query launch?"""
import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from nglod_b200 import ops
from helpers import rand5_model
dev = torch.device('cuda', 0)
net, args = rand5_model(dev)
view = net.net_view()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
x = torch.rand(n, 3, device=dev) * 2 - 1
for _ in range(3):
    ops.sdf_forward(view, 4, x)
    torch.cuda.synchronize()
    print("LAUNCH_END", flush=True)
