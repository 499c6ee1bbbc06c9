"""GPU-box diagnostic: per-row error of the CUDA encoder against the fp64 oracle for one golden
case, next to the error of a tf32-operand emulation of the same math (is a large row error inherent
to tf32 operands or a kernel bug?).  Test tooling, not product."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle import rrt_oracle as O
from golden_util import load_case
import gpu_util as G


def tf32_round(t):
    # round-to-nearest-even on the 13 dropped mantissa bits (cvt.rna rounds ties away; same scale)
    i = t.float().contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF
    return i.view(torch.float32).double()


def main(name):
    cfg, w, x, gold = load_case(name)
    ref = O.encoder_forward(x, w, cfg, "spec")
    # emulation: every matmul operand rounded to tf32, everything else fp64
    import torch.nn.functional as F
    orig_linear, orig_matmul = F.linear, torch.Tensor.__matmul__
    F.linear = lambda a, wt, b=None: orig_linear(tf32_round(a), tf32_round(wt), b)
    torch.Tensor.__matmul__ = lambda a, b: orig_matmul(tf32_round(a), tf32_round(b))
    try:
        emu = O.encoder_forward(x, w, cfg, "spec")
    finally:
        F.linear, torch.Tensor.__matmul__ = orig_linear, orig_matmul
    m = G.make_encoder(cfg, w)
    with torch.no_grad():
        y = m(x.float().cuda())
    y = y.cpu().double()
    rn = ref.norm(dim=1)
    e_gpu = (y - ref).norm(dim=1) / rn
    e_emu = (emu - ref).norm(dim=1) / rn
    print(f"{name}: overall rel gpu {O.rel_err(y, ref):.3e} emu {O.rel_err(emu, ref):.3e}")
    top = torch.topk(e_gpu, 8).indices.tolist()
    H, rs, _ = O.grid_geometry(x.shape[0], cfg.region_num, cfg.region_size, cfg.min_region_num, cfg.min_region_ratio)
    for t in top:
        print(f"  token {t:6d} (row {t // H}, col {t % H}) gpu {e_gpu[t]:.3e} emu {e_emu[t]:.3e} "
              f"|x| {x[t].norm():.2f}")
    print("  median row err gpu %.3e emu %.3e ; max emu %.3e" % (e_gpu.median(), e_emu.median(), e_emu.max()))
    # intermediate: R-MSA block only
    p = "layers.0."
    if cfg.n_layers > 1:
        r1 = x + O.rmsa_block(O.layer_norm(x, w[p + "norm.weight"], w[p + "norm.bias"]), w, p + "attn.", cfg, "spec")
        y1 = G.rmsa_block(m, 0, x.float().cuda()).cpu().double()
        e1 = (y1 - r1).norm(dim=1) / r1.norm(dim=1)
        print("  rmsa block: overall %.3e max row %.3e at token %d" % (O.rel_err(y1, r1), e1.max(), int(e1.argmax())))
        if cfg.cr_msa:
            r2 = r1 + O.crmsa_block(O.layer_norm(r1, w["cr_msa.norm.weight"], w["cr_msa.norm.bias"]), w, "cr_msa.attn.", cfg, "spec")
            y2 = G.crmsa_block(m, r1.float().cuda(), None, False).cpu().double()
            e2 = (y2 - r2).norm(dim=1) / r2.norm(dim=1)
            print("  crmsa block (exact x1 in): overall %.3e max row %.3e at token %d" % (O.rel_err(y2, r2), e2.max(), int(e2.argmax())))


if __name__ == "__main__":
    for n in sys.argv[1:] or ["d256_g4"]:
        main(n)
