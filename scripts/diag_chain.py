"""Bisect the backward chain of a p=0.3 step: gradients at every module boundary, gru impl tc vs fp32 (GEMMs fp32)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import mtad_gat_oracle as orc
from tests.golden_cases import inputs_for
from tests.test_gpu_parity import build, loss_fn
import mtad_gat_pytorch_b200 as mg
from mtad_gat_pytorch_b200 import functional as F

C2 = dict(n_features=38, window_size=100, out_dim=38, forecast_n_layers=3, dropout=0.3)
cfg = orc.Config(**C2)
params = orc.make_params(cfg, seed=70, dtype=np.float64)
B, S = 256, 424242
x, y = inputs_for(cfg, B, 70)
mg.set_gemm_impl("fp32")
via_model = len(sys.argv) > 1 and sys.argv[1] == "model"

def run(gru_impl):
    mg.set_gru_impl(gru_impl)
    m = build(C2, params, train=True)
    mg.manual_seed(S)
    xt = torch.from_numpy(x.astype(np.float32)).cuda().requires_grad_(True)
    yt = torch.from_numpy(y.astype(np.float32)).cuda()
    cap = {}
    def hook(name):
        return lambda g: cap.__setitem__(name, g.detach().clone())
    seed = F.fresh_seed(xt.device)
    for mod in m._seeded():
        mod._step_seed = seed
    xc = m.conv(xt); xc.register_hook(hook("d_xc(total)"))
    hf = m.feature_gat(xc); hf.register_hook(hook("d_hfeat"))
    ht = m.temporal_gat(xc); ht.register_hook(hook("d_htemp"))
    h_end = m.gru.forward_slices([xc, hf, ht]); h_end.register_hook(hook("d_h_end"))
    a = h_end
    fm = m.forecasting_model
    for i in range(len(fm.layers) - 1):
        a = F.LinearFn.apply(a, fm.layers[i].weight, fm.layers[i].bias, 1, 0.3, seed, F.RNG_MLP0 + i)
        a.register_hook(hook(f"d_mlp_a{i}"))
        cap[f"mlp_a{i}"] = a.detach().clone()
    preds = F.LinearFn.apply(a, fm.layers[-1].weight, fm.layers[-1].bias, 0, 0.0, None, 0)
    preds.register_hook(hook("d_preds"))
    recons = m.recon_model(h_end); recons.register_hook(hook("d_recons"))
    cap["preds"], cap["recons"], cap["h_end"] = preds.detach().clone(), recons.detach().clone(), h_end.detach().clone()
    loss_fn(xt, yt, preds, recons, None).backward()
    torch.cuda.synchronize()
    cap["dx"] = xt.grad.clone()
    return cap

a, b = run("fp32"), run("tc")
for k in ("h_end", "preds", "recons", "mlp_a0", "mlp_a1", "mlp_a2", "d_preds", "d_recons", "d_mlp_a2", "d_mlp_a1", "d_mlp_a0", "d_h_end", "d_hfeat", "d_htemp", "d_xc(total)", "dx"):
    d = (b[k] - a[k]).abs().reshape(B, -1).amax(1) / a[k].abs().max()
    bad = (d > 1e-3).nonzero().flatten().tolist()
    extra = ""
    if k.startswith("mlp_a"):
        extra = f" gate flips {int(((a[k] > 0) != (b[k] > 0)).sum())}"
    print(f"{k:14s} max rel err {float(d.max()):.1e} bad windows {bad}{extra}")
