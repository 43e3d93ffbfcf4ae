cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python scripts/diag_modules.py 256 0.3 2>&1 | grep "^\[" 
python scripts/diag_b256.py 256 2>&1 | grep -v Warn | grep "^\[\|dx err" | tail -12
timeout 300 python bench.py --steps 100 --warmup 5 --skip-cpu --skip-ref-cuda --sustain-s 0 > gpurun_out/r2_bench_precise.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2_bench_precise.log'):
    if l.startswith('{'):
        d=json.loads(l); print('value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'launches',d['gpu_launches'])
PY
timeout 900 python -m pytest tests/test_gpu_bench_config.py -q --timeout 600 -s > gpurun_out/r2_pytest_benchcfg.log 2>&1; echo "pytest benchcfg rc=$?"; grep -E "^\[|passed|failed|Error|error|assert" gpurun_out/r2_pytest_benchcfg.log | cut -c1-400 | tail -40
