cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_nccl_model.py -q --timeout 500 -s > gpurun_out/r2_pytest_nccl.log 2>&1; echo "pytest nccl rc=$?"; grep -E "^\[|passed|failed|Error|error|skipped" gpurun_out/r2_pytest_nccl.log | cut -c1-400 | tail -12
for v in "" "--no-overlap-comm" "--eager-comm"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 100 --warmup 5 --sustain-s 0 $v > gpurun_out/r2_bench_n2$v.log 2>&1; echo "bench n2 [$v] rc=$?"
  python - <<PY
import json
for l in open('gpurun_out/r2_bench_n2$v.log'):
    if l.startswith('{'):
        d=json.loads(l); print('N=2 [$v] value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'comm',d['config']['comm'])
PY
  tail -2 "gpurun_out/r2_bench_n2$v.log" | grep -v '^{' | cut -c1-300
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 100 --warmup 5 --sustain-s 0 --scaling strong > gpurun_out/r2_bench_n2_strong.log 2>&1; echo "bench n2 strong rc=$?"
grep '^{' gpurun_out/r2_bench_n2_strong.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=2 strong value',round(d['value']),'ms',round(d['ms_per_step'],4), d['config']['per_gpu_batch'])"
timeout 300 python bench.py --steps 100 --warmup 5 --sustain-s 0 --skip-cpu --skip-ref-cuda > gpurun_out/r2_bench_n1.log 2>&1
grep '^{' gpurun_out/r2_bench_n1.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=1 value',round(d['value']),'ms',round(d['ms_per_step'],4))"
timeout 900 python -m pytest tests/test_gpu_bench_config.py -q --timeout 600 -s -k "large_shape" > gpurun_out/r2_pytest_benchcfg2.log 2>&1; echo "pytest large rc=$?"; grep -E "^\[|passed|failed|Error|error" gpurun_out/r2_pytest_benchcfg2.log | cut -c1-400 | tail -12
