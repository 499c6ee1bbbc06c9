"""GPU-box tool: phase timeline (clock64) of the region-resident attention kernel at N=9000."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from rrt_mil_b200 import cabi, RRTEncoder
import gpu_util as G
NAMES = ["start", "issued", "landed", "qprime", "core", "end"] if os.environ.get("RRT_ATTN") != "tc05" else ["alloc", "qloaded", "toeplitz", "kv", "softmax0", "o0_sm1", "sync", "end"]
m = RRTEncoder(need_init=True).cuda().eval()
x = torch.randn(9000, 512, device="cuda")
tr = torch.zeros(64, 8, dtype=torch.int64, device="cuda")
with torch.no_grad():
    for _ in range(20): G.rmsa_block(m, 0, x)
    torch.cuda.synchronize()
    cabi.lib().rrt_debug_set_attn_trace(tr.data_ptr())
    G.rmsa_block(m, 0, x)
    torch.cuda.synchronize()
    cabi.lib().rrt_debug_set_attn_trace(None)
for c in range(2):
    t = tr[c].tolist()
    print(f"cta{c}: " + " ".join(f"{n}={t[i]-t[0]}" for i, n in enumerate(NAMES)))
if os.environ.get("RRT_ATTN") == "tc05":
    F = ["sm_enter", "s_ready", "pass1", "pfree", "pass2", "o_enter", "o_ready", "o_stored"]
    t0 = tr[0].tolist()[0]
    for row, who in ((8, "thread0/blk0"), (9, "thread128/blk1")):
        t = tr[row].tolist()
        print(f"  {who}: " + " ".join(f"{n}={t[i]-t0}" for i, n in enumerate(F)))
    sys.exit(0)

# fused CR-MSA landmarks kernel (same debug buffer)
NAMES2 = ["start", "setup", "pass1", "softmax", "pass2", "end"]
x1 = torch.randn(9000, 512, device="cuda")
tr.zero_()
with torch.no_grad():
    G.crmsa_block(m, x1, None, True)
    torch.cuda.synchronize()
    cabi.lib().rrt_debug_set_attn_trace(tr.data_ptr())
    G.crmsa_block(m, x1, None, True)
    torch.cuda.synchronize()
    cabi.lib().rrt_debug_set_attn_trace(None)
for c in range(3):
    t = tr[c].tolist()
    print(f"landmarks cta{c}: " + " ".join(f"{n}={t[i]-t[0]}" for i, n in enumerate(NAMES2)))

g0 = tr[:, 6].cpu(); g1 = tr[:, 7].cpu()
base = int(g0.min())
print("landmarks globaltimer (ns): CTA start offsets", sorted((g0 - base).tolist())[:6], "...", sorted((g0 - base).tolist())[-4:])
print("   CTA end offsets", sorted((g1 - base).tolist())[:4], "...", sorted((g1 - base).tolist())[-4:])
print("   per-CTA durations ns: min %d max %d" % (int((g1 - g0).min()), int((g1 - g0).max())))
