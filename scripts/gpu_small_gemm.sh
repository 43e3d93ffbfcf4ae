cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python scripts/small_gemm.py
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:tc_gemm -s 3 -c 1 -o gpurun_out/small_gemm python scripts/small_gemm.py > gpurun_out/small_gemm.log 2>&1; echo rc=$?
