cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_bench_config.py -q --timeout 150 -s -k "train_step_graph" > gpurun_out/r2_pytest_ts.log 2>&1; echo "pytest rc=$?"; grep -E "^\[trainstep|passed|failed" gpurun_out/r2_pytest_ts.log | cut -c1-200 | tail -12
for n in 1 2; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2965$n bench.py --gpus $n --config c4 --steps 10 --warmup 3 --sustain-s 0 --skip-cpu --skip-ref-cuda > gpurun_out/r2_bench_c4_n$n.log 2>&1; echo "bench c4 n=$n rc=$?"
grep '^{' gpurun_out/r2_bench_c4_n$n.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C4 N=$n', d['scaling'], 'value',round(d['value']),'ms',round(d['ms_per_step'],2),'per-gpu batch',d['config']['per_gpu_batch'])"
tail -2 gpurun_out/r2_bench_c4_n$n.log | grep -v '^{' | cut -c1-200
done
