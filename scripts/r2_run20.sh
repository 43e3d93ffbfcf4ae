cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -q --timeout 200 -x -s -k "ksplit_bptt" > gpurun_out/r2_pytest_ksplit.log 2>&1; echo "pytest ksplit rc=$?"; grep -E "^\[bptt|passed|failed|Error|error|trap|illegal|launch" gpurun_out/r2_pytest_ksplit.log | cut -c1-250 | tail -14
python - <<'PY'
import torch, sys
sys.path.insert(0,'.')
import mtad_gat_pytorch_b200 as mg
from mtad_gat_pytorch_b200 import kernel_bench
torch.manual_seed(0)
m = mg.MTAD_GAT(38,100,38,forecast_n_layers=3,dropout=0.3).cuda()
peaks={"hbm_gbs":6582.2,"bf16_tflops":1705.2,"bf16_tflops_sustained":1435.8}
flush = torch.empty(256*1024*1024//4, device='cuda')
for tag in ("unitsplit","ksplit"):
    mg.set_gru_bptt(tag)
    rows = kernel_bench.recurrence_rooflines(256, 100, 38, 150, m.gru.gru.weight_hh_l0.detach(), m.gru.gru.bias_hh_l0.detach(), peaks, torch.device('cuda'), flush, True)
    print(tag, [(r['kernel'], round(r['ms'],4)) for r in rows])
PY
timeout 300 python bench.py --steps 200 --warmup 5 --skip-cpu --skip-ref-cuda --sustain-s 0 > gpurun_out/r2_bench_ksplit.log 2>&1; echo "bench rc=$?"
grep '^{' gpurun_out/r2_bench_ksplit.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']))"
tail -3 gpurun_out/r2_bench_ksplit.log | grep -v '^{' | cut -c1-300
