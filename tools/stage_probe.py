"""GPU-box tool: stage timing of the CR-MSA block alone vs inside the full encoder."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from rrt_mil_b200 import cabi, RRTEncoder
import gpu_util as G
m = RRTEncoder(need_init=True).cuda().eval()
x = torch.randn(int(sys.argv[1]) if len(sys.argv) > 1 else 9000, 512, device="cuda")
def probe(tag, fn, n=10):
    with torch.no_grad():
        for _ in range(3): fn()
        torch.cuda.synchronize()
        cabi.stage_timing(True)
        for _ in range(n): fn()
        torch.cuda.synchronize()
        st = cabi.read_stage_timing(); cabi.stage_timing(False)
    print(tag, {k: round(v[0] / v[1] * 1e3, 1) for k, v in st.items()})
probe("crmsa block alone   ", lambda: G.crmsa_block(m, x, None, True))
probe("full encoder        ", lambda: m(x))
probe("rmsa block alone    ", lambda: G.rmsa_block(m, 0, x))
