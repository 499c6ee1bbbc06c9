"""GPU-box tool: where does an RRTMIL inference forward of one N~9000 bag spend its time (device and host)?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rrt_mil_b200 import RRTMIL, cabi
torch.manual_seed(0)
m = RRTMIL(input_dim=1024, n_classes=2).cuda().eval()
bags = [torch.randn(n, 1024, device="cuda") for n in (9000, 8200, 9900, 8700)]
def t_dev(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); h1 = time.perf_counter(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3, (h1 - h0) / n * 1e6
with torch.no_grad():
    d, h = t_dev(lambda: [m(b.unsqueeze(0)) for b in bags])
    print(f"RRTMIL forward: {d / 4:.1f} us/bag device, {h / 4:.1f} us/bag host enqueue")
    enc = m.online_encoder
    h0 = [torch.randn(b.shape[0], 512, device="cuda") for b in bags]
    d, h = t_dev(lambda: [enc(x) for x in h0])
    print(f"encoder alone : {d / 4:.1f} us/bag device, {h / 4:.1f} us/bag host enqueue")
    cabi.stage_timing(True)
    for _ in range(5): [m(b.unsqueeze(0)) for b in bags]
    torch.cuda.synchronize()
    st = cabi.read_stage_timing(); cabi.stage_timing(False)
    print({k: round(v[0] / 20 * 1e3, 1) for k, v in st.items()})
