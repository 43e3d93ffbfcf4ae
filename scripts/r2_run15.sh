cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for n in 2 3 4 6; do
MTADGAT_PARAM_STREAMS=$n timeout 300 python bench.py --steps 200 --warmup 5 --skip-cpu --skip-ref-cuda --sustain-s 0 > gpurun_out/r2_bench_ps$n.log 2>&1; echo "bench ps=$n rc=$?"
grep '^{' gpurun_out/r2_bench_ps$n.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ps=$n value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']))"
done
