set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "probe or tc_gemm" > gpurun_out/probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/probe.log
grep -E "tc probe|linear|passed|failed|Error" gpurun_out/probe.log | head -30
timeout 1500 python -m pytest tests -q -m gpu -s --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "worst|passed|failed|FAILED|rc=" gpurun_out/pytest_gpu.log | tail -50
timeout 600 python bench.py --steps 20 --warmup 3 --skip-cpu > gpurun_out/bench_graph.log 2>&1; echo "rc=$?" >> gpurun_out/bench_graph.log
python - <<'PY'
import json
for l in open('gpurun_out/bench_graph.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'])
        for k in d['kernels']: print(f"  {k['kernel']:14s} {k['ms']*1e3:9.1f} us  frac {k['frac']:.4f} ({k['bound']})")
PY
tail -3 gpurun_out/bench_graph.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/prof_step.py > gpurun_out/ncu_step.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections, re
with open('gpurun_out/launches.csv') as f:
    lines=[l for l in f if not l.startswith('==')]
r=csv.DictReader(lines)
agg=collections.OrderedDict()
for row in r:
    name=row['Kernel Name']; v=float(row['Metric Value'].replace(',',''))
    unit=row['Metric Unit']
    if unit=='ns': v/=1000.0
    elif unit=='ms': v*=1000.0
    name=re.sub(r'<unnamed>::','',name)[:100]
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
print('total us (3 steps)',tot)
for k,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:45]:
    print(f"{t/3:10.1f} us/step {c//3:4d}x  {100*t/tot:5.1f}%  {k}")
PY
