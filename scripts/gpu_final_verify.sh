cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_final.log 2>&1; echo "bench rc=$?"
grep '^{' gpurun_out/bench_final.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','roofline','clocks')}); print(d['cpu_baseline'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-300
timeout 300 python scripts/c3_throughput.py 2>&1 | tail -1
timeout 600 python scripts/timeline.py 0 gpurun_out/r1_timeline_graph_step.json 2>&1 | tail -7
