"""GPU-box tool (ncu target): N full RRTMIL train steps (BASELINE configs[4] shape: 1024 -> 512, epeg_k 21, crmsa_k 5,
one N=9000 bag; forward, cross entropy, backward, fused Adam), nothing else."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rrt_mil_b200 import RRTMIL
from rrt_mil_b200.optim import Adam
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
torch.manual_seed(2021)
tm = RRTMIL(input_dim=1024, n_classes=2, epeg_k=21, crmsa_k=5, n_layers=2).cuda().train()
opt = Adam(tm.parameters(), lr=2e-4, weight_decay=1e-5)
tb = torch.randn(1, 9000, 1024, device="cuda")
label = torch.tensor([1], device="cuda")
for _ in range(n):
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.functional.cross_entropy(tm(tb), label)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print("ok", float(loss))
