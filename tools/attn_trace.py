"""GPU-box tool: phase timeline (clock64) of the region-resident attention kernel at N=9000."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from rrt_mil_b200 import cabi, RRTEncoder
import gpu_util as G
NAMES = ["start", "issued", "landed", "qprime", "core", "end"]
m = RRTEncoder(need_init=True).cuda().eval()
x = torch.randn(9000, 512, device="cuda")
tr = torch.zeros(8, 8, dtype=torch.int64, device="cuda")
with torch.no_grad():
    G.rmsa_block(m, 0, x); G.rmsa_block(m, 0, x)
    torch.cuda.synchronize()
    cabi.lib().rrt_debug_set_attn_trace(tr.data_ptr())
    G.rmsa_block(m, 0, x)
    torch.cuda.synchronize()
    cabi.lib().rrt_debug_set_attn_trace(None)
for c in range(4):
    t = tr[c].tolist()
    print(f"cta{c}: " + " ".join(f"{n}={t[i]-t[0]}" for i, n in enumerate(NAMES)))
