cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gat_fused_win -s 8 -c 4 -o gpurun_out/r2_gat_fused -f python scripts/prof_gat.py > gpurun_out/r2_ncu_gat.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2_ncu_gat.log
ls -la gpurun_out/r2_gat_fused.ncu-rep
