cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bench_config.py -q --timeout 300 -x -k "adam or train_step or first_call or reseed or pipelined" > gpurun_out/r2_pytest_adam.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_adam.log | cut -c1-300
timeout 300 python bench.py --steps 200 --warmup 5 --skip-cpu --skip-ref-cuda --sustain-s 0 > gpurun_out/r2_bench_adam.log 2>&1; echo "bench rc=$?"
grep '^{' gpurun_out/r2_bench_adam.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'launches/step',d['gpu_launches']/d['steps'])"
tail -3 gpurun_out/r2_bench_adam.log | grep -v '^{' | cut -c1-300
