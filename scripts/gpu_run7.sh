set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 120 python scripts/gru_phases.py 2>&1 | tail -2 | cut -c1-400
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "golden_forward_backward and tc" > gpurun_out/tc_golden.log 2>&1; echo "rc=$?" >> gpurun_out/tc_golden.log
grep -E "worst|passed|failed|rc=|Error" gpurun_out/tc_golden.log | grep -v tc1 | head -12
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --skip-cpu > gpurun_out/bench_graph.log 2>&1; echo "rc=$?" >> gpurun_out/bench_graph.log
python - <<'PY'
import json
for l in open('gpurun_out/bench_graph.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'])
        for k in d['kernels']: print(f"  {k['kernel']:14s} {k['ms']*1e3:9.1f} us  frac {k['frac']:.4f} ({k['bound']})")
PY
