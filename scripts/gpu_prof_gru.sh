set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gru_cl -s 4 -c 2 -o gpurun_out/prof_gru_cl python scripts/prof_step.py > gpurun_out/prof_gru.log 2>&1; echo "rc=$?"
tail -3 gpurun_out/prof_gru.log
