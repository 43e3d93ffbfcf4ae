"""v1 tiny layer: split vs fused vs the oracle (which one is off in lin.bias?)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import mtad_gat_oracle as orc
import mtad_gat_pytorch_b200 as mg
torch.manual_seed(11)
k, n, B = 6, 20, 37
x = torch.rand(B, n, k, device="cuda"); go = torch.randn(B, n, k, device="cuda")
for cls, feat in ((mg.FeatureAttentionLayer, True), (mg.TemporalAttentionLayer, False)):
    layer = cls(k, n, 0.0, 0.2, None, False).cuda().eval()
    with torch.no_grad(): layer.bias.normal_()
    p = {nm: q.detach().cpu().numpy().astype(np.float64) for nm, q in layer.named_parameters()}
    out, cache = orc.gat_fwd(x.cpu().numpy().astype(np.float64), p["lin.weight"], p["lin.bias"], p["a"], p["bias"], 0.2, feat, False, None)
    dx, dw, db, da, dbias = orc.gat_bwd(go.cpu().numpy().astype(np.float64), cache)
    ref = {"lin.weight": dw, "lin.bias": db, "a": da, "bias": dbias}
    for impl in ("split", "fused", "split", "fused"):
        mg.set_gat_impl(impl)
        layer.zero_grad(set_to_none=True)
        xi = x.clone().requires_grad_(True)
        y = layer(xi); y.backward(go); torch.cuda.synchronize()
        errs = {nm: float(np.abs(q.grad.cpu().numpy() - ref[nm]).max() / np.abs(ref[nm]).max()) for nm, q in layer.named_parameters()}
        errs["out"] = float(np.abs(y.detach().cpu().numpy() - out).max()); errs["dx"] = float(np.abs(xi.grad.cpu().numpy() - dx).max() / np.abs(dx).max())
        print(cls.__name__, impl, {a: f"{b:.1e}" for a, b in errs.items()}, "| |db| max", float(np.abs(db).max()))
mg.set_gat_impl("fused")
