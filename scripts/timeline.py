"""Kernel timeline of one CUDA-graph replay of the training step (CUPTI activity records via torch.profiler):
start offset, duration, stream of every kernel -> gpurun_out/timeline.csv, plus a per-stream summary."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mtad_gat_pytorch_b200 as mg
from mtad_gat_pytorch_b200 import training as mgt
from torch.profiler import profile, ProfilerActivity
split = int(sys.argv[1]) if len(sys.argv) > 1 else 0
if split: mg.set_gru_split(split)
pipes = int(sys.argv[3]) if len(sys.argv) > 3 else 1
B = 256
torch.manual_seed(0)
m = mg.MTAD_GAT(38, 100, 38, forecast_n_layers=3, dropout=0.3).cuda().train()
opt = torch.optim.Adam(m.parameters(), lr=1e-3, capturable=True, fused=True)
step = mgt.TrainStep(m, opt, batch=B, use_graph=True, world_size=1, pipeline=pipes)
x = torch.rand(B, 100, 38, device="cuda"); y = torch.rand(B, 1, 38, device="cuda")
for _ in range(5): step.run_device(x, y)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step.run_device(x, y)
        torch.cuda.synchronize()
out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/timeline.json"
prof.export_chrome_trace(out)
ev = [e for e in json.load(open(out))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
# last replay = last third
n = len(ev) // 3
last = ev[2 * n:]
t0 = last[0]["ts"]
rows = []
for e in last:
    rows.append((e["ts"] - t0, e["dur"], e["args"].get("stream"), e["name"][:90]))
with open(out.replace(".json", ".csv"), "w") as f:
    f.write("start_us,dur_us,stream,kernel\n")
    for r in rows: f.write("%.1f,%.1f,%s,%s\n" % r)
end = max(r[0] + r[1] for r in rows)
print("kernels", len(rows), "span_us", round(end, 1))
bys = {}
for r in rows: bys.setdefault(r[2], []).append(r)
for s, rs in bys.items():
    print("stream", s, "n", len(rs), "busy_us", round(sum(r[1] for r in rs), 1), "first", round(rs[0][0], 1), "last_end", round(max(r[0] + r[1] for r in rs), 1))
os.remove(out)
