"""Per-phase cycle counters of the cluster GRU forward kernel (CTA 0), encoder and decoder."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mtad_gat_pytorch_b200 as mg
from mtad_gat_pytorch_b200._lib import lib
dbg = torch.zeros(16, dtype=torch.int64, device="cuda")
lib.mtadgat_gru_debug_buffer(dbg.data_ptr())
torch.manual_seed(0)
m = mg.MTAD_GAT(38, 100, 38, forecast_n_layers=3, dropout=0.3).cuda().train()
m.branch_parallel = False
x = torch.rand(256, 100, 38, device="cuda")
names = ["mma.wait_h", "mma.issue+commit", "-", "epi.wait_acc", "epi.drain+bar1", "epi.math", "epi.st_async", "epi.tail(stores+prefetch)"]
if len(sys.argv) > 1:
    mg.set_gru_split(int(sys.argv[1]))
for trial in range(2):
    p, r = m(x)
    torch.cuda.synchronize()
    v = dbg.cpu().tolist()
    print("last GRU launch (decoder) phases, cycles/step:", {k: v[i] for i, k in enumerate(names)}, "sum epi", sum(v[2:9]), "sum mma", v[0] + v[1])
