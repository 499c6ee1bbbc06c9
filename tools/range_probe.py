"""GPU-box tool (ncu --replay-mode range target): ONE 16-bag step of the throughput mode (8 lanes) between
cudaProfilerStart / Stop, so that ncu reports the DRAM / L2 bytes of the whole concurrent step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rrt_mil_b200 import RRTEncoder
lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 8
m = RRTEncoder(need_init=True).cuda().eval()
bags = [torch.randn(9000, 512, device="cuda") for _ in range(16)]
outs = [torch.empty_like(b) for b in bags]
with torch.no_grad():
    for _ in range(3):
        m.forward_bags(bags, outs, lanes=lanes)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    m.forward_bags(bags, outs, lanes=lanes)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("ok")
