"""tcgen05 micro-measurements: cost of a recurrence step's MMA chain; accumulator lane map of an M=64 MMA."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mtad_gat_pytorch_b200._lib import lib, check
st = torch.cuda.current_stream().cuda_stream
out = torch.zeros(1, dtype=torch.int64, device="cuda")
for (nt, kc, M, N, rs, ni, mode) in [(6, 10, 128, 16, 64, 1, 0), (6, 10, 128, 16, 64, 1, 1), (6, 10, 128, 16, 64, 2, 0), (6, 10, 128, 16, 64, 2, 1),
                                     (6, 10, 128, 16, 64, 3, 1), (6, 10, 128, 16, 64, 6, 0), (6, 10, 128, 16, 64, 6, 1), (2, 29, 128, 16, 24, 2, 1),
                                     (8, 8, 128, 16, 24, 8, 1), (8, 8, 128, 16, 24, 4, 1), (1, 10, 128, 16, 128, 1, 1), (1, 29, 128, 16, 128, 1, 1)]:
    check(lib.mtadgat_tc_mma_bench(nt, kc, M, N, 200, rs, ni, mode, out.data_ptr(), st))
    torch.cuda.synchronize()
    c = int(out.item())
    print(f"tiles={nt} kchunks={kc} M={M} N={N} issuers={ni} mode={mode}: {c} cycles/step  ({c / (nt * kc):.1f} per MMA)")
# M=64 lane map
g = torch.Generator().manual_seed(0)
A = torch.randn(64, 32, generator=g).cuda(); Bm = torch.randn(16, 32, generator=g).cuda()
D = torch.full((128, 16), float("nan"), device="cuda")
check(lib.mtadgat_tc_probe(A.data_ptr(), Bm.data_ptr(), D.data_ptr(), 64, 0, 32, 16, 0, 64, st))
torch.cuda.synchronize()
ref = A.half().float() @ Bm.half().float().t()
lanes = []
for r in range(64):
    d = (D - ref[r]).abs().max(dim=1).values
    lanes.append(int(torch.argmin(d)))
print("M=64: row -> TMEM lane:", lanes)
