cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q --timeout 300 -x -k "golden or linear or fused or variants or smd" > gpurun_out/r2_pytest_pack.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_pack.log | cut -c1-300
timeout 300 python bench.py --steps 200 --warmup 5 --skip-cpu --skip-ref-cuda --sustain-s 0 > gpurun_out/r2_bench_pack.log 2>&1; echo "bench rc=$?"
grep '^{' gpurun_out/r2_bench_pack.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'traffic',d['roofline']['traffic'],d['roofline']['traffic_src'])"
timeout 200 python scripts/timeline.py 0 gpurun_out/r2_timeline_pack.json 1 2>&1 | grep -v Warn | tail -2
grep -E "pack_kernel" gpurun_out/r2_timeline_pack.csv | cut -c1-80 | head -40
