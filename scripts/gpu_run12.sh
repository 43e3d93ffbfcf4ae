cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python scripts/small_gemm.py
timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python scripts/timeline.py 2 gpurun_out/timeline_g3.json 2>&1 | tail -6
timeout 600 python bench.py --steps 30 --warmup 3 --skip-cpu > gpurun_out/bench_graph.log 2>&1; echo "rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/bench_graph.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'])
PY
tail -3 gpurun_out/bench_graph.log | grep -v '^{' | cut -c1-300
python - <<'PY'
# fixed vs per-step cost of the recurrence launches
import torch, sys, os
sys.path.insert(0, os.getcwd())
import mtad_gat_pytorch_b200 as mg
from mtad_gat_pytorch_b200 import kernel_bench
peaks={"hbm_gbs":6582.2,"bf16_tflops":1435.8}
flush = torch.empty(256*1024*1024//4, device="cuda")
torch.manual_seed(0)
H=150
w=torch.randn(3*H,H,device="cuda")*0.08; b=torch.randn(3*H,device="cuda")*0.1
for n in (100,200,400):
    r=kernel_bench.recurrence_rooflines(256,n,H,w,b,peaks,"cuda",flush)
    print(n,[ (x['kernel'],round(x['ms']*1e3,1)) for x in r])
PY
