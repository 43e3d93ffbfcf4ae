set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -s --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "bwd\]|passed|failed|FAILED|rc=" gpurun_out/pytest_gpu.log | tail -12
timeout 600 python bench.py --steps 30 --warmup 3 --skip-cpu > gpurun_out/bench_graph.log 2>&1; echo "rc=$?" >> gpurun_out/bench_graph.log
python - <<'PY'
import json
for l in open('gpurun_out/bench_graph.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e'],'launches',d['gpu_launches'])
PY
tail -2 gpurun_out/bench_graph.log | cut -c1-200
timeout 300 python scripts/c3_throughput.py 2>&1 | tail -2
