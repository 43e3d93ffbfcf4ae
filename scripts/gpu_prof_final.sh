cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_final.log 2>&1; echo "bench rc=$?"
timeout 600 python scripts/timeline.py 0 gpurun_out/r1_timeline_graph_step.json 2>&1 | tail -6
# launch list of the bench command (eager launches so that every kernel is a separate ncu record)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r1_final_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-graph --skip-cpu > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gru_cl_(fwd|bwd)_kernel" -s 4 -c 4 -o gpurun_out/r1_final_gru_cl python scripts/prof_step.py > gpurun_out/prof_gru.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gat_(score_win|bwd1|bwd2_win)_kernel" -s 6 -c 6 -o gpurun_out/r1_final_gat python scripts/prof_step.py > gpurun_out/prof_gat.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"gemm2_kernel|pack_kernel" -s 60 -c 8 -o gpurun_out/r1_final_gemm python scripts/prof_step.py > gpurun_out/prof_gemm.log 2>&1; echo "rc=$?"
