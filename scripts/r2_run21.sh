cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -4 gpurun_out/r2_pytest_gpu.log | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_driverlike.log 2>&1; echo "bench rc=$?"
grep '^{' gpurun_out/r2_bench_driverlike.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'cpu',d['cpu_baseline']['value'],'traffic',d['roofline']['traffic'],'clocks',d['clocks'])"
