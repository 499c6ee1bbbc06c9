"""GPU-box tool: host enqueue time vs GPU time of forward_bags (is the launch path the bottleneck?)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rrt_mil_b200 import RRTEncoder
m = RRTEncoder(need_init=True).cuda().eval()
bags = [torch.randn(9000, 512, device="cuda") for _ in range(16)]
outs = [torch.empty_like(b) for b in bags]
with torch.no_grad():
    for lanes in (1, 4):
        for _ in range(5): m.forward_bags(bags, outs, lanes=lanes)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10): m.forward_bags(bags, outs, lanes=lanes)
        t_host = time.perf_counter() - t0
        torch.cuda.synchronize()
        t_all = time.perf_counter() - t0
        print(f"lanes={lanes}: host enqueue {t_host/160*1e6:.1f} us/bag, total {t_all/160*1e6:.1f} us/bag")
