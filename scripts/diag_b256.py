"""Diagnostic: where does the B=256 tc train-mode parity error come from?  mode x dropout x batch grid vs the oracle."""
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
grid = [(m, p, B) for B in (int(a) for a in sys.argv[1].split(",")) for p in (0.0, 0.3) for m in ("fp32", "tc")]
cache = {}
for mode, p, B in grid:
    mg.set_mode(mode)
    x, y = inputs_for(cfg, B, 70)
    m = build(C2, params, train=p > 0)
    S = 424242
    mg.manual_seed(S)
    masks = ot.masks_for_seed(ot.seed_after(S, 1), cfg, B, 0.3) if p > 0 else None
    xt = torch.from_numpy(x.astype(np.float32)).cuda().requires_grad_(True)
    yt = torch.from_numpy(y.astype(np.float32)).cuda()
    preds, recons = m(xt)
    loss = loss_fn(xt, yt, preds, recons, None)
    loss.backward()
    torch.cuda.synchronize()
    key = (p, B)
    if key not in cache:
        t0 = time.time()
        cache[key] = ot.loss_fwd_bwd_chunked(x, y, params, cfg, masks=masks, chunk=32)
        print(f"  oracle B={B} p={p}: {time.time()-t0:.1f}s", flush=True)
    l_ref, _, _, p_ref, r_ref, dx_ref, g_ref = cache[key]
    errs = {"preds": rel(preds, p_ref), "recons": rel(recons, r_ref), "dx": rel(xt.grad, dx_ref)}
    for pname, q in m.named_parameters():
        errs["g." + pname] = rel(q.grad, g_ref[pname])
    top = sorted(errs.items(), key=lambda kv: -kv[1])[:8]
    print(f"[{mode} p={p} B={B}] " + " ".join(f"{k}={v:.1e}" for k, v in top), flush=True)
    # localise the dx error over windows
    d = np.abs(xt.grad.cpu().numpy() - dx_ref).reshape(B, -1).max(1) / np.abs(dx_ref).max()
    print("    dx err by window block of 16:", " ".join(f"{d[i:i+16].max():.0e}" for i in range(0, B, 16)), flush=True)
