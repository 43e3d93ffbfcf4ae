cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 100 --warmup 5 --sustain-s 0 > gpurun_out/r2_bench_n2.log 2>&1; echo "bench n2 rc=$?"
grep '^{' gpurun_out/r2_bench_n2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=2 value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']), d['config']['comm'])"
tail -2 gpurun_out/r2_bench_n2.log | grep -v '^{' | cut -c1-300
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29623 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r2_bench_ref_n2.log 2>&1; echo "ref n2 rc=$?"; grep '^{' gpurun_out/r2_bench_ref_n2.log | cut -c1-200
timeout 280 python -m pytest tests/test_nccl_model.py -q --timeout 270 -s > gpurun_out/r2_pytest_nccl.log 2>&1; echo "pytest nccl rc=$?"; grep -E "^\[|passed|failed|Error|error|skipped" gpurun_out/r2_pytest_nccl.log | cut -c1-500 | tail -8
