set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# (1) launch list of the bench command itself (eager launches so that every kernel is visible to ncu)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r1_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-graph --skip-cpu > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"
# (2) full capture of the dominant kernels: cluster GRU BPTT + forward, GAT score forward, GAT bwd2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gru_cl_(fwd|bwd)_kernel" -s 4 -c 4 -o gpurun_out/r1_full_gru_cl python scripts/prof_step.py > gpurun_out/prof_gru.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gat_(score_fwd|bwd1|bwd2)_kernel" -s 6 -c 6 -o gpurun_out/r1_full_gat python scripts/prof_step.py > gpurun_out/prof_gat.log 2>&1; echo "rc=$?"
tail -2 gpurun_out/prof_gru.log gpurun_out/prof_gat.log
