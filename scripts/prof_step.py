"""One training step at SMD shape (B=256), for ncu captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mtad_gat_pytorch_b200 as mg
from mtad_gat_pytorch_b200.training import rmse_losses
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(0)
m = mg.MTAD_GAT(38, 100, 38, forecast_n_layers=3, dropout=0.3).cuda().train()
x = torch.rand(B, 100, 38, device="cuda"); y = torch.rand(B, 1, 38, device="cuda")
for _ in range(3):
    m.zero_grad(set_to_none=True)
    p, r = m(x)
    fl, rl = rmse_losses(x, y, p, r)
    (fl + rl).backward()
torch.cuda.synchronize()
print("done", float(fl + rl))
