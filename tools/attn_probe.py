"""GPU-box tool: the tcgen05 attention core against the mma.sync core -- R-MSA block parity over a range of
shapes, stage time of each, and the clock64 phase trace of CTA 0 (softmax thread 0, per item).

    python tools/attn_probe.py [--trace] [--quick]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402

import gpu_util as G  # noqa: E402
from rrt_mil_b200 import RRTEncoder, cabi  # noqa: E402

lib = cabi.lib()
CASES = [
    (9000, dict()), (512, dict()), (63, dict()), (65, dict()), (1, dict()), (9216, dict()),
    (1300, dict(region_num=4, epeg_k=9)), (3000, dict(region_num=16, epeg_k=21)),
    (20000, dict(region_num=16)), (50000, dict(region_num=16)), (9000, dict(epeg=False)),
    (16000, dict(region_num=8)),            # P = 256 with EPEG: not supported by the tcgen05 core -> mma.sync
    (14000, dict(region_num=8, epeg_k=9)),  # P = 225
    (700, dict(mlp_dim=256, n_heads=4, crmsa_heads=4)),
]
if "--quick" in sys.argv:
    CASES = CASES[:3]


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    cabi.stage_timing(True)
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    st = cabi.read_stage_timing()
    cabi.stage_timing(False)
    return st["rmsa_attention"][0] / st["rmsa_attention"][1] * 1e3


bad = 0
with torch.no_grad():
    for L, over in CASES:
        torch.manual_seed(L)
        m = RRTEncoder(need_init=True, **over).cuda().eval()
        for q in m.parameters():   # non-trivial biases / taps
            if q.dim() == 1:
                q.add_(0.1 * torch.randn_like(q))
        x = torch.randn(L, m.final_dim, device="cuda")
        lib.rrt_debug_set_attention_kernel(0)
        ref = G.rmsa_block(m, 0, x)
        t0 = timed(lambda: G.rmsa_block(m, 0, x))
        lib.rrt_debug_set_attention_kernel(2)
        got = G.rmsa_block(m, 0, x)
        got2 = G.rmsa_block(m, 0, x)
        t1 = timed(lambda: G.rmsa_block(m, 0, x))
        torch.cuda.synchronize()
        r = rel(got, ref)
        ok = r < 2e-4 and torch.equal(got, got2) and bool(torch.isfinite(got).all())
        bad += not ok
        print(f"L={L:6d} {over}: rel(tc05, mma)={r:.2e} deterministic={torch.equal(got, got2)} "
              f"attention us: mma {t0:.1f}  tc05 {t1:.1f}  {'OK' if ok else 'FAIL'}", flush=True)

    if "--trace" in sys.argv:
        m = RRTEncoder(need_init=True).cuda().eval()
        x = torch.randn(9000, 512, device="cuda")
        tr = torch.zeros(64, 8, dtype=torch.int64, device="cuda")
        for _ in range(5):
            G.rmsa_block(m, 0, x)
        torch.cuda.synchronize()
        lib.rrt_debug_set_attn_trace(tr.data_ptr())
        G.rmsa_block(m, 0, x)
        torch.cuda.synchronize()
        lib.rrt_debug_set_attn_trace(None)
        names = ["item", "s_full", "pass1", "p_ready", "o_full", "stored"]
        t = tr.cpu().view(-1)[:56].view(7, 8)
        t0 = int(t[6, 6])
        print(f"  kernel entry=0 after setup/pdl_wait={int(t[6, 7]) - t0}  (thread 0 = softmax warp 0: its items are 0, 2, 4, ...)")
        for n in range(3):
            print(f"  item {2 * n}: " + " ".join(f"{nm}={int(t[n, i]) - t0}" for i, nm in enumerate(names)))
        full = tr.cpu()
        for stp in range(6):
            r = full[8 + stp].tolist()
            print(f"  helper0 step {stp}: " + " ".join(f"{nm}={r[i] - t0 if r[i] else -1}" for i, nm in
                  enumerate(["start", "tail_ready", "tail_done", "epeg_ready", "epeg_done"])))
        for n in range(4):
            r = full[16 + n].tolist()
            print(f"  issuer item {n}: " + " ".join(f"{nm}={r[i] - t0 if r[i] else -1}" for i, nm in
                  enumerate(["top", "p_ready", "o_issued", "s_go", "s_issued", "loads_done"])))
        for n in range(2):
            r = full[24 + n].tolist()
            print(f"  issuer first steps j={n}: " + " ".join(f"{nm}={r[i] - t0 if r[i] else -1}" for i, nm in
                  enumerate(["s_section", "qp_full", "kv_full", "s_issued"])))
lib.rrt_debug_set_attention_kernel(1)
print("FAILED" if bad else "all OK")
sys.exit(1 if bad else 0)
