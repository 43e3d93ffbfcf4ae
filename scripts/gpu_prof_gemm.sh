set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"DghTT|Cat3AT" -s 2 -c 2 -o gpurun_out/prof_gemm python scripts/prof_step.py > gpurun_out/prof_gemm.log 2>&1; echo "rc=$?"
tail -3 gpurun_out/prof_gemm.log
