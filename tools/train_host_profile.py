"""GPU-box tool: where does the HOST time of one RRTMIL train step go?  cProfile over N steps enqueued into an
empty queue (synchronise between steps), sorted by own time and by cumulative time.  Not a bench number."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rrt_mil_b200 import RRTMIL, cabi  # noqa: E402
from rrt_mil_b200.optim import Adam  # noqa: E402

torch.manual_seed(2021)
dev = torch.device("cuda:0")
tm = RRTMIL(input_dim=1024, n_classes=2, epeg_k=21, crmsa_k=5, n_layers=2).to(dev).train()
opt = Adam(tm.parameters(), lr=2e-4, weight_decay=1e-5)
tb = torch.randn(1, 9000, 1024, device=dev)
label = torch.tensor([1], device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.functional.cross_entropy(tm(tb), label)
    loss.backward()
    opt.step()
    return loss


for _ in range(5):
    step()
torch.cuda.synchronize()
n = 30
host = []
for _ in range(n):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step()
    host.append((time.perf_counter() - t0) * 1e6)
torch.cuda.synchronize()
host.sort()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
l0 = cabi.launch_count()
e0.record()
for _ in range(n):
    step()
e1.record()
torch.cuda.synchronize()
print(f"host enqueue median {host[n // 2]:.0f} us / step, min {host[0]:.0f}; back-to-back {e0.elapsed_time(e1) / n * 1e3:.0f} us / step; "
      f"{(cabi.launch_count() - l0) // n} launches counted by the library")
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    torch.cuda.synchronize()
    step()
pr.disable()
torch.cuda.synchronize()
for key in ("tottime", "cumtime"):
    print(f"---- top by {key} (totals over {n} steps; divide by {n}) ----")
    pstats.Stats(pr, stream=sys.stdout).strip_dirs().sort_stats(key).print_stats(28)
