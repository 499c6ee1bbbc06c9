"""GPU-box tool (ncu target): N forwards of one N=9000 bag through the default encoder, nothing else."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rrt_mil_b200 import RRTEncoder
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
L = int(sys.argv[2]) if len(sys.argv) > 2 else 9000
m = RRTEncoder(need_init=True).cuda().eval()
x = torch.randn(L, 512, device="cuda")
with torch.no_grad():
    for _ in range(n):
        y = m(x)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
