cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bench_config.py -q --timeout 600 -s > gpurun_out/r2_pytest_benchcfg.log 2>&1; echo "pytest benchcfg rc=$?"; grep -E "^\[|passed|failed|Error|error" gpurun_out/r2_pytest_benchcfg.log | cut -c1-400 | tail -30
timeout 200 python scripts/timeline.py 0 gpurun_out/r2_timeline_p1.json 1 2>&1 | grep -v Warn | tail -8
timeout 200 python scripts/timeline.py 0 gpurun_out/r2_timeline_p2.json 2 2>&1 | grep -v Warn | tail -8
for c in c3 c4 c5; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 3 --skip-cpu --skip-ref-cuda --sustain-s 0 > gpurun_out/r2_bench_$c.log 2>&1; echo "bench $c rc=$?"
  python - <<PY
import json
for l in open('gpurun_out/r2_bench_$c.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$c value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'launches',d['gpu_launches'], 'roof', d['roofline']['kernel'], round(d['roofline']['frac'],4))
PY
  tail -2 gpurun_out/r2_bench_$c.log | grep -v '^{' | cut -c1-300
done
