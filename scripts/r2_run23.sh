cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 4 --steps 100 --warmup 5 --sustain-s 0 > gpurun_out/r2_bench_n4.log 2>&1; echo "bench n4 rc=$?"
grep '^{' gpurun_out/r2_bench_n4.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=4 value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']))"
tail -2 gpurun_out/r2_bench_n4.log | grep -v '^{' | cut -c1-200
