cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/bench_ddp2.log 2>&1; echo "rc=$?"
grep '^{' gpurun_out/bench_ddp2.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('value',d['value'],'ms',d['ms_per_step'],'n',d['n_gpus'],'e2e',d['e2e']['value'])"
tail -3 gpurun_out/bench_ddp2.log | grep -v '^{' | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -2 | cut -c1-400
