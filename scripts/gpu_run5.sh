set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python scripts/mma_bench.py > gpurun_out/mma_bench.log 2>&1; cat gpurun_out/mma_bench.log
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --skip-cpu > gpurun_out/bench_graph.log 2>&1; echo "rc=$?" >> gpurun_out/bench_graph.log
python - <<'PY'
import json
for l in open('gpurun_out/bench_graph.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'])
PY
tail -2 gpurun_out/bench_graph.log | cut -c1-300
