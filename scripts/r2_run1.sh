cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
for p in 1 2; do
  timeout 300 python bench.py --steps 100 --warmup 5 --skip-cpu --skip-ref-cuda --pipeline $p --sustain-s 0 > gpurun_out/r2_bench_p$p.log 2>&1; echo "bench pipeline=$p rc=$?"
done
python - <<'PY'
import json
for p in (1,2):
    for l in open(f'gpurun_out/r2_bench_p{p}.log'):
        if l.startswith('{'):
            d=json.loads(l); print('pipeline',p,'value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'launches',d['gpu_launches'])
PY
tail -3 gpurun_out/r2_bench_p2.log | grep -v '^{' | cut -c1-400
timeout 900 python -m pytest tests/test_gpu_bench_config.py -q -x --timeout 600 -s > gpurun_out/r2_pytest_benchcfg.log 2>&1; echo "pytest benchcfg rc=$?"; grep -E "^\[|passed|failed|Error|error" gpurun_out/r2_pytest_benchcfg.log | cut -c1-300 | tail -30
timeout 900 python -m pytest tests -q -m gpu --timeout 600 --deselect tests/test_gpu_bench_config.py > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -5 gpurun_out/r2_pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r2_bench_full.log 2>&1; echo "bench full rc=$?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.log 2>&1; echo "bench ref rc=$?"
python - <<'PY'
import json
for f in ('r2_bench_full','r2_bench_ref'):
    for l in open(f'gpurun_out/{f}.log'):
        if l.startswith('{'):
            d=json.loads(l); d.pop('kernels',None); print(f, json.dumps(d)[:2500])
PY
