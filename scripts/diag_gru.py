"""Isolate the tc-recurrence backward error: replay the encoder GRU and the decoder on the REAL tensors of a p=0.3 step
(inputs and incoming gradients captured from the full model), tc vs fp32 recurrence."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import mtad_gat_oracle as orc
from tests.golden_cases import inputs_for
from tests.test_gpu_parity import build, loss_fn
import mtad_gat_pytorch_b200 as mg

C2 = dict(n_features=38, window_size=100, out_dim=38, forecast_n_layers=3, dropout=0.3)
cfg = orc.Config(**C2)
params = orc.make_params(cfg, seed=70, dtype=np.float64)
B, S = 256, 424242
x, y = inputs_for(cfg, B, 70)
mg.set_mode("fp32")
m = build(C2, params, train=True)
mg.manual_seed(S)
xt = torch.from_numpy(x.astype(np.float32)).cuda()
yt = torch.from_numpy(y.astype(np.float32)).cuda()
cap = {}
xc = m.conv(xt)
for mod in (m.feature_gat, m.temporal_gat):
    mod._step_seed = None
hf = m.feature_gat(xc); ht = m.temporal_gat(xc)
h_end = m.gru.forward_slices([xc, hf, ht])
h_end.register_hook(lambda g: cap.__setitem__("dh_end", g.clone()))
preds = m.forecasting_model(h_end)
recons = m.recon_model(h_end)
recons.register_hook(lambda g: cap.__setitem__("drecons", g.clone()))
loss_fn(xt, yt, preds, recons, None).backward()
torch.cuda.synchronize()
xc, hf, ht, h_end = xc.detach(), hf.detach(), ht.detach(), h_end.detach()
print("dh_end absmax", float(cap["dh_end"].abs().max()), "per-window max (top5):", torch.topk(cap["dh_end"].abs().amax(1), 5))
print("drecons absmax", float(cap["drecons"].abs().max()))

def rel(a, b): return float((a - b).abs().max() / b.abs().max())

def enc(impl):
    mg.set_gru_impl(impl)
    ins = [t.clone().requires_grad_(True) for t in (xc, hf, ht)]
    for p in m.gru.parameters(): p.grad = None
    h = m.gru.forward_slices(ins)
    h.backward(cap["dh_end"])
    torch.cuda.synchronize()
    return h.detach(), [t.grad.clone() for t in ins], [p.grad.clone() for p in m.gru.parameters()]

def dec(impl):
    mg.set_gru_impl(impl)
    hi = h_end.clone().requires_grad_(True)
    for p in m.recon_model.parameters(): p.grad = None
    r = m.recon_model(hi)
    r.backward(cap["drecons"])
    torch.cuda.synchronize()
    return r.detach(), [hi.grad.clone()], [p.grad.clone() for p in m.recon_model.parameters()]

for name, fn in (("encoder", enc), ("decoder", dec)):
    a, b = fn("fp32"), fn("tc")
    e_out = rel(b[0], a[0])
    e_in = [rel(u, v) for u, v in zip(b[1], a[1])]
    e_p = [rel(u, v) for u, v in zip(b[2], a[2])]
    d = (b[1][0] - a[1][0]).abs().reshape(B, -1).amax(1) / a[1][0].abs().max()
    bad = (d > 1e-3).nonzero().flatten().tolist()
    print(f"[{name}] tc vs fp32: out {e_out:.1e} input grads {['%.1e' % e for e in e_in]} params {['%.1e' % e for e in e_p]} bad windows {bad}")
    if name == "encoder" and bad:
        w = bad[0]
        dd = (b[1][0][w] - a[1][0][w]).abs().amax(1) / a[1][0].abs().max()       # by time step
        print("   window", w, "err by t (every 10):", ["%.0e" % float(v) for v in dd[::10]], "|dh_end| of window", float(cap["dh_end"][w].abs().max()))
