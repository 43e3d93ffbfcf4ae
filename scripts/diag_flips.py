"""Which implementation choice produces the window-local gradient blips at B=256, p=0.3?  (a) gemm/gru impl matrix vs the
oracle, (b) count ReLU gate differences of the MLP hidden layers between implementations (same masks)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import mtad_gat_oracle as orc
from tests import oracle_tools as ot
from tests.golden_cases import inputs_for
from tests.test_gpu_parity import build, loss_fn, rel
import mtad_gat_pytorch_b200 as mg

C2 = dict(n_features=38, window_size=100, out_dim=38, forecast_n_layers=3, dropout=0.3)
cfg = orc.Config(**C2)
params = orc.make_params(cfg, seed=70, dtype=np.float64)
B, S = 256, 424242
x, y = inputs_for(cfg, B, 70)
masks = ot.masks_for_seed(ot.seed_after(S, 1), cfg, B, 0.3)
ref = ot.loss_fwd_bwd_chunked(x, y, params, cfg, masks=masks, chunk=32)
l_ref, _, _, p_ref, r_ref, dx_ref, g_ref = ref
gates = {}
for gemm, gru in (("fp32", "fp32"), ("tc", "fp32"), ("fp32", "tc"), ("tc", "tc")):
    mg.set_gemm_impl(gemm); mg.set_gru_impl(gru)
    m = build(C2, params, train=True)
    acts = []
    hooks = [l.register_forward_hook(lambda mod, i, o: None) for l in []]
    mg.manual_seed(S)
    xt = torch.from_numpy(x.astype(np.float32)).cuda().requires_grad_(True)
    yt = torch.from_numpy(y.astype(np.float32)).cuda()
    # capture MLP hidden activations by re-running the head on h_end
    preds, recons = m(xt)
    loss_fn(xt, yt, preds, recons, None).backward()
    torch.cuda.synchronize()
    errs = {"preds": rel(preds, p_ref), "recons": rel(recons, r_ref), "dx": rel(xt.grad, dx_ref)}
    for pname, q in m.named_parameters():
        errs["g." + pname] = rel(q.grad, g_ref[pname])
    top = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
    d = np.abs(xt.grad.cpu().numpy() - dx_ref).reshape(B, -1).max(1) / np.abs(dx_ref).max()
    print(f"[gemm={gemm} gru={gru}] " + " ".join(f"{k}={v:.1e}" for k, v in top) + f" | windows dx>1e-3: {np.nonzero(d > 1e-3)[0].tolist()}", flush=True)
    # gate bits of the MLP: eval-mode (no dropout) hidden activations from the encoder state
    m.eval()
    with torch.no_grad():
        xc = m.conv(xt.detach()); hf = m.feature_gat(xc); ht = m.temporal_gat(xc)
        h = m.gru.forward_slices([xc, hf, ht])
        g = []
        a = h
        from mtad_gat_pytorch_b200 import functional as F
        for i in range(len(m.forecasting_model.layers) - 1):
            L = m.forecasting_model.layers[i]
            a = F.LinearFn.apply(a, L.weight, L.bias, 1, 0.0, None, 0)
            g.append((a > 0).cpu().numpy())
        gates[(gemm, gru)] = (g, (xc > 0).cpu().numpy(), h.cpu().numpy())
base = gates[("fp32", "fp32")]
for key, (g, cg, h) in gates.items():
    diffs = [int((a != b).sum()) for a, b in zip(g, base[0])]
    rows = sorted(set(np.nonzero((g[0] != base[0][0]).any(1))[0].tolist()))
    print(f"{key}: MLP gate flips per layer vs fp32/fp32 (eval, no dropout): {diffs}; layer-0 rows {rows}; conv gate flips {int((cg != base[1]).sum())}; "
          f"max |h_end diff| {np.abs(h - base[2]).max():.1e}")
