"""Bisect: each sub-module alone at B=256, train mode p=0.3, same seed, tc vs fp32 (fp32 is exact to ~1e-6)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mtad_gat_pytorch_b200 as mg

torch.manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
k = int(sys.argv[3]) if len(sys.argv) > 3 else 38
n = int(sys.argv[4]) if len(sys.argv) > 4 else 100
H = 150
PD = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
model = mg.MTAD_GAT(k, n, k, forecast_n_layers=3, dropout=PD).cuda().train()
with torch.no_grad():
    model.feature_gat.bias.normal_(); model.temporal_gat.bias.normal_()

def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))

def run(name, fn, params, x, mode, seed=5):
    mg.set_mode(mode); mg.manual_seed(seed)
    for p in params: p.grad = None
    xi = x.clone().requires_grad_(True)
    out = fn(xi)
    g = torch.Generator(device="cuda").manual_seed(7)
    go = torch.randn(out.shape, device="cuda", generator=g)
    out.backward(go)
    torch.cuda.synchronize()
    return out.detach().clone(), xi.grad.clone(), [p.grad.clone() for p in params]

x = torch.rand(B, n, k, device="cuda")
h = torch.randn(B, H, device="cuda") * 0.5
cases = [
    ("conv", lambda t: model.conv(t), list(model.conv.parameters()), x),
    ("feature_gat", lambda t: model.feature_gat(t), list(model.feature_gat.parameters()), x),
    ("temporal_gat", lambda t: model.temporal_gat(t), list(model.temporal_gat.parameters()), x),
    ("gru", lambda t: model.gru.forward_slices([t, t * 0.5, t * 0.25]), list(model.gru.parameters()), x),
    ("mlp", lambda t: model.forecasting_model(t), list(model.forecasting_model.parameters()), h),
    ("recon", lambda t: model.recon_model(t), list(model.recon_model.parameters()), h),
]
for name, fn, params, inp in cases:
    ref = run(name, fn, params, inp, "fp32")
    a = run(name, fn, params, inp, "tc")
    b = run(name, fn, params, inp, "tc")
    names = [n_ for n_, _ in [(q, 0) for q in range(len(params))]]
    e_out, e_dx = rel(a[0], ref[0]), rel(a[1], ref[1])
    e_p = [rel(u, v) for u, v in zip(a[2], ref[2])]
    rep = (rel(b[0], a[0]), rel(b[1], a[1]), max(rel(u, v) for u, v in zip(b[2], a[2])))
    d = (a[1] - ref[1]).abs().reshape(B, -1).max(1).values / ref[1].abs().max()
    bad = (d > 1e-3).nonzero().flatten().tolist()
    print(f"[{name}] tc-vs-fp32: out {e_out:.1e} dx {e_dx:.1e} params {' '.join('%.1e' % e for e in e_p)} | tc repeat: out {rep[0]:.1e} dx {rep[1]:.1e} p {rep[2]:.1e} | windows with dx err>1e-3: {bad[:12]}", flush=True)
