cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scoring.py -q --timeout 300 -x -k "golden or variants or score_series or smd" > gpurun_out/r2_pytest_rep.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_pytest_rep.log | cut -c1-300
timeout 300 python bench.py --steps 200 --warmup 5 --skip-cpu --skip-ref-cuda --sustain-s 0 > gpurun_out/r2_bench_rep.log 2>&1; echo "bench rc=$?"
grep '^{' gpurun_out/r2_bench_rep.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']))"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --steps 100 --warmup 5 --sustain-s 0 > gpurun_out/r2_bench_n2.log 2>&1; echo "bench n2 rc=$?"
grep '^{' gpurun_out/r2_bench_n2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=2 value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']))"
tail -2 gpurun_out/r2_bench_n2.log | grep -v '^{' | cut -c1-300
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29633 bench.py --gpus 2 --steps 100 --warmup 5 --sustain-s 0 --scaling strong > gpurun_out/r2_bench_n2_strong.log 2>&1; echo "bench n2 strong rc=$?"
grep '^{' gpurun_out/r2_bench_n2_strong.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=2 strong value',round(d['value']),'ms',round(d['ms_per_step'],4),'per-gpu batch',d['config']['per_gpu_batch'])"
timeout 150 python -m pytest tests/test_nccl_model.py -q --timeout 140 > gpurun_out/r2_pytest_nccl.log 2>&1; echo "pytest nccl rc=$?"; tail -1 gpurun_out/r2_pytest_nccl.log
