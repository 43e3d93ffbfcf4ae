cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scoring.py -q --timeout 600 -s > gpurun_out/r2_pytest_scoring.log 2>&1; echo "pytest scoring rc=$?"; grep -E "^\[|passed|failed|Error|error" gpurun_out/r2_pytest_scoring.log | cut -c1-300 | tail -30
timeout 900 python -m pytest tests/test_gpu_bench_config.py -q --timeout 600 -s -k "large_shape or pipelined" > gpurun_out/r2_pytest_benchcfg2.log 2>&1; echo "pytest benchcfg rc=$?"; grep -E "^\[|passed|failed|Error|error" gpurun_out/r2_pytest_benchcfg2.log | cut -c1-300 | tail -12
timeout 600 python bench.py --config c3 --steps 20 --warmup 3 --skip-cpu --skip-ref-cuda --sustain-s 0 > gpurun_out/r2_bench_c3.log 2>&1; echo "bench c3 rc=$?"
python - <<PY
import json
for l in open('gpurun_out/r2_bench_c3.log'):
    if l.startswith('{'):
        d=json.loads(l); print('c3 value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']), d['single_pass_scoring'])
PY
tail -3 gpurun_out/r2_bench_c3.log | grep -v '^{' | cut -c1-300
