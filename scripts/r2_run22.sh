cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for sp in 4 1; do
timeout 200 python bench.py --steps 200 --warmup 5 --skip-cpu --skip-ref-cuda --sustain-s 0 --gru-split $sp > gpurun_out/r2_bench_split$sp.log 2>&1; echo "bench split=$sp rc=$?"
grep '^{' gpurun_out/r2_bench_split$sp.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('split $sp value',round(d['value']),'ms',round(d['ms_per_step'],4), [ (k['kernel'],round(k['ms'],4)) for k in d['kernels'] if 'recurrence' in k['kernel']])"
tail -2 gpurun_out/r2_bench_split$sp.log | grep -v '^{' | cut -c1-200
done
