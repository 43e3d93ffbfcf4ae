cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_scoring.py -q --timeout 200 -k "epsilon" > gpurun_out/r2_pytest_eps.log 2>&1; echo "pytest eps rc=$?"; tail -2 gpurun_out/r2_pytest_eps.log | cut -c1-200
timeout 300 python bench.py --steps 200 --warmup 5 --skip-cpu --skip-ref-cuda --sustain-s 0 > gpurun_out/r2_bench_zw.log 2>&1; echo "bench rc=$?"
grep '^{' gpurun_out/r2_bench_zw.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']))"
timeout 200 python scripts/timeline.py 0 gpurun_out/r2_timeline_zw.json 1 2>&1 | grep -v Warn | tail -1
grep -E "gru_cl|absmax|rep_dh|zero_word" gpurun_out/r2_timeline_zw.csv | cut -c1-70
