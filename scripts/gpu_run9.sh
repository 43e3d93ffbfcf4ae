set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "split or gru or golden or smd or variants" --timeout 600 2>&1 | tail -5
for sp in 1 2 4; do
timeout 300 python scripts/gru_phases.py $sp 2>&1 | tail -1
timeout 600 python bench.py --steps 30 --warmup 3 --skip-cpu --gru-split $sp > gpurun_out/bench_split$sp.log 2>&1; echo "rc=$?"
python - <<PY
import json
for l in open('gpurun_out/bench_split$sp.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('split $sp value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
        for k in d['kernels']:
            if 'gru' in k['kernel'] or 'recon' in k['kernel']: print('  ',k['kernel'],round(k['ms'],4),k['bound'],round(k['frac'],4))
PY
tail -3 gpurun_out/bench_split$sp.log | cut -c1-300 | grep -v '^{'
done
