set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gru_tc -s 4 -c 4 -o gpurun_out/prof_gru_tc python scripts/prof_step.py > gpurun_out/prof_gru.log 2>&1; echo "rc=$?"
tail -5 gpurun_out/prof_gru.log
ls -la gpurun_out/
