"""GPU-box tool: what each stage of the encoder costs while several bags are in flight.

Per-kernel CUDA-event intervals overlap when bags run on concurrent lanes, so they cannot be summed.
Instead: time the 16-bag step with every stage on, then with ONE stage's kernels not launched
(rrt_debug_skip_stages; results are wrong by construction, only the clock matters).  The drop is the
stage's marginal cost in throughput mode.  Also prints host enqueue time per bag.
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rrt_mil_b200 import RRTEncoder, cabi

lib = cabi.lib()
m = RRTEncoder(need_init=True).cuda().eval()
bags = [torch.randn(9000, 512, device="cuda") for _ in range(16)]
outs = [torch.empty_like(b) for b in bags]
names = [lib.rrt_stage_name(i).decode() for i in range(lib.rrt_stage_count())]


def timed(lanes, steps=20):
    with torch.no_grad():
        for _ in range(3):
            m.forward_bags(bags, outs, lanes=lanes)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            m.forward_bags(bags, outs, lanes=lanes)
        e1.record()
        t_host = time.perf_counter() - t0
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (steps * len(bags)), t_host * 1e6 / (steps * len(bags))


quick = "--quick" in sys.argv
print("RRT_GEMM_CLUSTER =", os.environ.get("RRT_GEMM_CLUSTER"), " RRT_ATTN =", os.environ.get("RRT_ATTN"))
for lanes in ((1, 4) if quick else ((8,) if "--l8" in sys.argv else (1, 2, 3, 4, 6, 8))):
    us, host = timed(lanes)
    print(f"lanes={lanes}: {us:7.2f} us/bag   host enqueue {host:6.2f} us/bag", flush=True)

fwd = ["ln_partition", "qkv_gemm", "rmsa_attention", "proj_gemm_residual", "crmsa_landmarks",
       "landmark_qkv_gemm", "landmark_attention", "landmark_proj_gemm", "crmsa_dispatch_final_ln"]
for lanes in (() if quick else ((8,) if "--l8" in sys.argv else (1, 4))):
    base, _ = timed(lanes)
    print(f"--- lanes={lanes}: all stages {base:.2f} us/bag")
    for n in fwd:
        lib.rrt_debug_skip_stages(1 << names.index(n))
        us, _ = timed(lanes)
        print(f"  without {n:26s} {us:7.2f} us/bag   marginal {base - us:6.2f}", flush=True)
    lib.rrt_debug_skip_stages(0)
