"""GPU-box tool: phase timeline (clock64) of the tcgen05 GEMM for the path's shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from rrt_mil_b200 import cabi, RRTEncoder
import gpu_util as G

NAMES = ["start", "setup", "tma0", "tmaN", "opnd0", "mmaN", "acc0", "epi0", "epiN", "end", "c0ld", "c0tr", "c0st", "c1ld", "c1tr", "c1st"]


def show(tag, tr, launch=0):
    tr = tr.cpu()[launch]
    for cta in range(2):
        t = tr[cta].tolist()
        base = t[0]
        print(f"  {tag} cta{cta}: " + " ".join(f"{n}={t[i]-base if t[i] else -1}" for i, n in enumerate(NAMES)))


def main():
    lib = cabi.lib()
    tr = torch.zeros(8, 8, 16, dtype=torch.int64, device="cuda")
    for (M, N, K) in [(9216, 1536, 512), (9216, 512, 512), (192, 1536, 512)]:
        a = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") / K ** 0.5
        b = torch.randn(N, device="cuda")
        G.linear_f16(a, w, b)
        torch.cuda.synchronize()
        tr.zero_()
        lib.rrt_debug_set_gemm_trace(tr.data_ptr())
        G.linear_f16(a, w, b)
        torch.cuda.synchronize()
        lib.rrt_debug_set_gemm_trace(None)
        for l in range(8):
            if int(tr[l, 0, 9]) != 0:
                show(f"linear {M}x{N}x{K} (fp32 out)", tr, l)
    # the proj GEMM inside the R-MSA block (residual-scatter epilogue): last GEMM of the block
    m = RRTEncoder(need_init=True).cuda().eval()
    x = torch.randn(9000, 512, device="cuda")
    with torch.no_grad():
        m(x); m(x)
        torch.cuda.synchronize()
        tr.zero_()
        lib.rrt_debug_set_gemm_trace(tr.data_ptr())
        G.rmsa_block(m, 0, x)
        torch.cuda.synchronize()
        lib.rrt_debug_set_gemm_trace(None)
    for l in range(8):
        if int(tr[l, 0, 9]) != 0:
            show(f"rmsa block gemm (launch slot {l}: qkv f16-out first, then proj)", tr, l)


if __name__ == "__main__":
    main()
