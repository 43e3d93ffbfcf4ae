cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -6 gpurun_out/r2_pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/r2_bench_full.log 2>&1; echo "bench full rc=$?"
timeout 500 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.log 2>&1; echo "bench ref rc=$?"
python - <<'PY'
import json
for f in ('r2_bench_full','r2_bench_ref'):
    for l in open(f'gpurun_out/{f}.log'):
        if l.startswith('{'):
            d=json.loads(l); d.pop('kernels',None); print(f, json.dumps(d)[:2200])
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gru_cl -s 8 -c 4 -o gpurun_out/r2_gru_cl -f python scripts/prof_step.py 256 > gpurun_out/r2_ncu_gru.log 2>&1; echo "ncu gru rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-graph --skip-cpu --skip-ref-cuda --sustain-s 0 > gpurun_out/r2_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
for c in c3 c4 c5; do
  timeout 500 python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/r2_bench_$c.log 2>&1; echo "bench $c rc=$?"
done
timeout 300 python bench.py --gatv1 --steps 100 --warmup 5 --skip-cpu --skip-ref-cuda > gpurun_out/r2_bench_gatv1.log 2>&1; echo "bench gatv1 rc=$?"
timeout 200 python scripts/timeline.py 0 gpurun_out/r2_timeline_graph_step.json 1 2>&1 | grep -v Warn | tail -3
