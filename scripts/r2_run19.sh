cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -4 gpurun_out/r2_pytest_gpu.log | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/r2_bench_full.log 2>&1; echo "bench full rc=$?"
timeout 500 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.log 2>&1; echo "bench ref rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-graph --skip-cpu --skip-ref-cuda --sustain-s 0 > gpurun_out/r2_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 500 python bench.py --config c3 --steps 20 --warmup 3 > gpurun_out/r2_bench_c3.log 2>&1; echo "bench c3 rc=$?"
timeout 300 python bench.py --gatv1 --steps 100 --warmup 5 --skip-cpu --skip-ref-cuda > gpurun_out/r2_bench_gatv1.log 2>&1; echo "bench gatv1 rc=$?"
timeout 200 python scripts/timeline.py 0 gpurun_out/r2_timeline_graph_step.json 1 2>&1 | grep -v Warn | tail -1
python - <<'PY'
import json
for f in ('r2_bench_full','r2_bench_ref','r2_bench_c3','r2_bench_gatv1'):
    for l in open(f'gpurun_out/{f}.log'):
        if l.startswith('{'):
            d=json.loads(l); print(f,'value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'refcuda', (d.get('reference_cuda') or {}).get('value'), 'roof', d.get('roofline',{}).get('frac'), d.get('roofline',{}).get('traffic'), (d.get('single_pass_scoring') or {}).get('timestamps_per_s'))
PY
