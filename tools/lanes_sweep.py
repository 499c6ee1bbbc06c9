"""GPU-box tool: us/bag of the 16-bag step for the lanes given on the command line (env knobs apply)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rrt_mil_b200 import RRTEncoder
m = RRTEncoder(need_init=True).cuda().eval()
bags = [torch.randn(9000, 512, device="cuda") for _ in range(16)]
outs = [torch.empty_like(b) for b in bags]
res = []
with torch.no_grad():
    for lanes in [int(a) for a in sys.argv[1:]] or [4]:
        for _ in range(3):
            m.forward_bags(bags, outs, lanes=lanes)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            m.forward_bags(bags, outs, lanes=lanes)
        e1.record()
        torch.cuda.synchronize()
        res.append(f"lanes={lanes}: {e0.elapsed_time(e1) * 1e3 / 320:.2f}")
print(" | ".join(res), flush=True)
