"""GPU-box tool: A/B of the experimental LayerNorm-fused QKV GEMM (resident A tile)
against the default ln_partition + QKV GEMM pair.

    python tools/fused_qkv_probe.py            # parity vs the default path, stage times, us/bag per lane count
    python tools/fused_qkv_probe.py --trace    # + clock64 phase stamps of the fused kernel (2 CTAs)
    python tools/fused_qkv_probe.py --pair     # the CTA-pair form (cta_group::2) instead of the single-CTA kernel

Stamps of the fused kernel: start, setup, tma0 (first W tile issued), tmaN = A tile filled (MMA side),
opnd0 (first W tile landed), mmaN (last MMA committed), acc0 (first accumulator ready), epi0 / epiN (first /
last tile stored), end.
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from rrt_mil_b200 import RRTEncoder, cabi
import gpu_util as G

lib = cabi.lib()
m = RRTEncoder(need_init=True).cuda().eval()
bags = [torch.randn(9000, 512, device="cuda") for _ in range(16)]
outs = [torch.empty_like(b) for b in bags]
NAMES = ["start", "setup", "tma0", "afill", "opnd0", "mmaN", "acc0", "epi0", "epiN", "end"]


MODE = 4 if "--pair" in sys.argv else 3   # 3: single-CTA fused kernel, 4: its CTA-pair form


def fused(on):
    lib.rrt_debug_set_gemm_cluster(MODE if on else 30)


def us_per_bag(lanes, steps=20):
    with torch.no_grad():
        for _ in range(3):
            m.forward_bags(bags, outs, lanes=lanes)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            m.forward_bags(bags, outs, lanes=lanes)
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (steps * len(bags))


def stage_table():
    cabi.stage_timing(True)
    with torch.no_grad():
        for _ in range(5):
            m.forward_bags(bags, outs, lanes=1)
    torch.cuda.synchronize()
    st = cabi.read_stage_timing()
    cabi.stage_timing(False)
    return {k: v[0] / v[1] * 1e3 for k, v in st.items()}


with torch.no_grad():
    x = bags[0]
    fused(False)
    ref_blk, ref = G.rmsa_block(m, 0, x), m(x)
    fused(True)
    got_blk, got = G.rmsa_block(m, 0, x), m(x)
    torch.cuda.synchronize()
    rel = lambda a, b: float((a - b).norm() / b.norm())
    print(f"parity vs default path: rmsa block rel {rel(got_blk, ref_blk):.2e}, encoder rel {rel(got, ref):.2e}")

for on in (False, True):
    fused(on)
    t = stage_table()
    print(("fused  " if on else "default"), " ".join(f"{k}={v:.1f}" for k, v in t.items() if "gemm" in k or "ln_" in k))
    print("        us/bag:", " ".join(f"lanes={l}: {us_per_bag(l):.2f}" for l in (1, 2, 4, 8)), flush=True)

if "--trace" in sys.argv:
    fused(True)
    tr = torch.zeros(8, 8, 16, dtype=torch.int64, device="cuda")
    with torch.no_grad():
        G.rmsa_block(m, 0, x)
        torch.cuda.synchronize()
        lib.rrt_debug_set_gemm_trace(tr.data_ptr())
        G.rmsa_block(m, 0, x)
        torch.cuda.synchronize()
        lib.rrt_debug_set_gemm_trace(None)
    t = tr.cpu()[0]  # launch slot 0 = the fused kernel (slot 1 = proj)
    for cta in range(2):
        row = t[cta].tolist()
        print(f"  fused cta{cta}: " + " ".join(f"{n}={row[i] - row[0] if row[i] else -1}" for i, n in enumerate(NAMES)))
fused(False)
