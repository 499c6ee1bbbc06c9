"""GPU-box tool: device time of the path on every BASELINE.json config shape (the bench line is configs[1];
the others are parity-test cases, timed here for DESIGN.md).  CUDA events, inputs resident, 3 warm-ups."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rrt_mil_b200 import RRTEncoder, RRTMIL


def timed(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


torch.manual_seed(2021)
rows = []
with torch.no_grad():
    enc = RRTEncoder(need_init=True).cuda().eval()
    for name, N in (("C1 N=512 fwd, one bag", 512), ("C2 N=9000 fwd, one bag", 9000)):
        x = torch.randn(N, 512, device="cuda")
        us = timed(lambda: enc(x))
        rows.append((name, us, N / us))
    bags = [torch.randn(9000, 512, device="cuda") for _ in range(16)]
    outs = [torch.empty_like(b) for b in bags]
    us = timed(lambda: enc.forward_bags(bags, outs, lanes=4), n=10) / 16
    rows.append(("C2 N=9000 fwd, 16 bags, 4 in flight (the bench line)", us, 9000 / us))
    enc16 = RRTEncoder(need_init=True, region_num=16).cuda().eval()
    x = torch.randn(50000, 512, device="cuda")
    us = timed(lambda: enc16(x), n=10)
    rows.append(("C4 N=50000 region_num=16 fwd, one bag", us, 50000 / us))
    b4 = [torch.randn(50000, 512, device="cuda") for _ in range(4)]
    o4 = [torch.empty_like(b) for b in b4]
    us = timed(lambda: enc16.forward_bags(b4, o4, lanes=4), n=5) / 4
    rows.append(("C4 N=50000 region_num=16 fwd, 4 bags in flight", us, 50000 / us))
    mil = RRTMIL(input_dim=1024, n_classes=2).cuda().eval()
    g = torch.Generator().manual_seed(3)
    lens = torch.randint(8000, 10001, (8,), generator=g).tolist()
    mb = [torch.randn(1, n, 1024, device="cuda") for n in lens]
    us = timed(lambda: [mil(b) for b in mb], n=10) / 8
    rows.append((f"C3 RRTMIL(1024->512 + encoder + DAttention head), 8 ragged bags N~U[8000,10000], per bag", us,
                 sum(lens) / 8 / us))
enc5 = RRTEncoder(need_init=True, epeg_k=21, crmsa_k=5).cuda().train()
x = torch.randn(9000, 512, device="cuda", requires_grad=True)
gout = torch.randn(9000, 512, device="cuda")
params = list(enc5.parameters())


def step():
    for p in params:
        p.grad = None
    x.grad = None
    enc5(x).backward(gout)


us = timed(step, n=10)
rows.append(("C5 encoder shape (epeg_k=21, crmsa_k=5) N=9000 train fwd+bwd (dropout 0.1), one bag", us, 9000 / us))
import torch.nn.functional as F
from rrt_mil_b200.optim import Adam
mil5 = RRTMIL(input_dim=1024, n_classes=2, epeg_k=21, crmsa_k=5).cuda().train()
opt = Adam(mil5.parameters(), lr=2e-4, weight_decay=1e-5)
bag = torch.randn(1, 9000, 1024, device="cuda")
lab = torch.tensor([1], device="cuda")


def mil_step():
    opt.zero_grad(set_to_none=True)
    F.cross_entropy(mil5(bag), lab).backward()
    opt.step()


us = timed(mil_step, n=10)
rows.append(("C5 RRTMIL(1024, epeg_k=21, crmsa_k=5) full train step: fwd + CE + bwd + Adam, N=9000, one bag", us, 9000 / us))
for name, us, mps in rows:
    print(f"{name:100s} {us:9.1f} us   {mps:8.2f} M patches/s")
