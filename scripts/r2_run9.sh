cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q --timeout 200 -x -k "fused_gat or golden or variants or extreme" > gpurun_out/r2_pytest_fused.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_fused.log | cut -c1-300
python - <<PY
import torch, time, sys
sys.path.insert(0,'.')
import mtad_gat_pytorch_b200 as mg
torch.manual_seed(0)
m = mg.MTAD_GAT(38,100,38,forecast_n_layers=3,dropout=0.3).cuda().eval()
x = torch.rand(256,100,38,device='cuda')
flush = torch.empty(64*1024*1024, device='cuda')
for layer in (m.feature_gat, m.temporal_gat):
    for mode in ("eval","train"):
        layer.train(mode=="train")
        ts=[]
        for _ in range(12):
            flush.zero_()
            xi = x.clone().requires_grad_(mode=="train")
            e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
            with torch.no_grad() if mode=="eval" else torch.enable_grad():
                e0.record()
                y=layer(xi)
            e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ts.sort()
        print("fused", type(layer).__name__, mode, "median ms", round(ts[len(ts)//2],4))
PY
timeout 300 python bench.py --steps 100 --warmup 5 --skip-cpu --skip-ref-cuda --sustain-s 0 > gpurun_out/r2_bench_fused.log 2>&1; echo "bench rc=$?"
grep '^{' gpurun_out/r2_bench_fused.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',round(d['value']),'ms',round(d['ms_per_step'],4),'launches',d['gpu_launches'])"
timeout 200 python scripts/timeline.py 0 gpurun_out/r2_timeline_fused.json 1 2>&1 | grep -v Warn | tail -3
