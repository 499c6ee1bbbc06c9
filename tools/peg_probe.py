"""GPU-box tool: time the PEG / PPEG kernel at N=9000, D=512 (and serve as the ncu target)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rrt_mil_b200 import cabi
lib = cabi.lib()
L, D = 9000, 512
x = torch.randn(L, D, device="cuda"); out = torch.empty_like(x)
ws = [torch.randn(D, 1, k, k, device="cuda") * 0.1 for k in (7, 5, 3)]
bs = [torch.randn(D, device="cuda") * 0.1 for _ in range(3)]
wp = (C.c_void_p * 3)(*[t.data_ptr() for t in ws]); bp = (C.c_void_p * 3)(*[t.data_ptr() for t in bs])
st = torch.cuda.current_stream().cuda_stream
for ppeg in (1, 0):
    for _ in range(3):
        cabi.check(lib.rrt_peg_forward(x.data_ptr(), out.data_ptr(), L, D, 7, ppeg, 0, wp, bp, st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        cabi.check(lib.rrt_peg_forward(x.data_ptr(), out.data_ptr(), L, D, 7, ppeg, 0, wp, bp, st))
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    print(f"{'PPEG' if ppeg else 'PEG'} k=7 N={L} D={D}: {us:.1f} us per call (fold + apply), algorithmic 2 x {L*D*4/1e6:.1f} MB -> {2*L*D*4/us/1e3:.0f} GB/s")
