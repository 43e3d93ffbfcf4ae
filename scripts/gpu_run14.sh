cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python scripts/gru_phases.py 2 2>&1 | tail -1
timeout 600 python bench.py --steps 30 --warmup 3 --skip-cpu > gpurun_out/bench_graph.log 2>&1; echo "rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/bench_graph.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'])
        for k in d['kernels']:
            if 'recurrence' in k['kernel']: print('  ',k['kernel'],round(k['ms'],4),k['bound'],round(k['frac'],4))
PY
tail -3 gpurun_out/bench_graph.log | grep -v '^{' | cut -c1-300
