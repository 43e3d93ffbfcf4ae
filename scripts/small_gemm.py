"""Tiny nn.Linear-shaped GEMM (M=256, I=150, O=150) straight through the C ABI: timing + target for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mtad_gat_pytorch_b200 as mg
from mtad_gat_pytorch_b200._lib import lib, check
torch.manual_seed(0)
M, I, O = 256, 150, 150
x = torch.randn(M, I, device="cuda"); w = torch.randn(O, I, device="cuda") * 0.1; b = torch.randn(O, device="cuda")
y = torch.empty(M, O, device="cuda"); seed = torch.zeros(1, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for act, p in ((0, 0.0), (1, 0.3)):
    f = lambda: check(lib.mtadgat_linear_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), M, I, O, act, p, seed.data_ptr(), 16, st))
    for _ in range(3): f()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    print("act", act, "p", p, "us", sorted(ts)[len(ts) // 2], flush=True)
