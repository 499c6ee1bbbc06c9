"""GPU-box tool: single-bag latency, eager C call vs CUDA-graph replay (rrt_mil_b200/graph.py, EXPERIMENTAL)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rrt_mil_b200 import RRTEncoder
from rrt_mil_b200.graph import GraphedForward

m = RRTEncoder(need_init=True).cuda().eval()


def timed(fn, n=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


with torch.no_grad():
    for lengths, lanes in [([512], 1), ([2000], 1), ([9000], 1), ([9000] * 4, 4), ([9000] * 16, 8)]:
        bags = [torch.randn(n, 512, device="cuda") for n in lengths]
        outs = [torch.empty_like(b) for b in bags]
        g = GraphedForward(m, lengths, lanes=lanes)
        for b, buf in zip(bags, g.inputs):
            buf.copy_(b)
        eager = timed(lambda: m.forward_bags(bags, outs, lanes=lanes))
        graph = timed(g.replay)
        same = all(torch.equal(a, b) for a, b in zip(outs, g.outputs))
        print(f"bags={lengths[0]}x{len(lengths)} lanes={lanes}: eager {eager / len(lengths):7.2f} us/bag, "
              f"graph {graph / len(lengths):7.2f} us/bag, identical={same}", flush=True)
