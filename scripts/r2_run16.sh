cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q --timeout 300 -x -k "golden or fused or variants or dropout or accumulation" > gpurun_out/r2_pytest_late.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_pytest_late.log | cut -c1-300
timeout 300 python bench.py --steps 200 --warmup 5 --skip-cpu --skip-ref-cuda --sustain-s 0 > gpurun_out/r2_bench_late.log 2>&1; echo "bench rc=$?"
grep '^{' gpurun_out/r2_bench_late.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']))"
timeout 200 python scripts/timeline.py 0 gpurun_out/r2_timeline_late.json 1 2>&1 | grep -v Warn | tail -1
tail -32 gpurun_out/r2_timeline_late.csv | cut -c1-75
