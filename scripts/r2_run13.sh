cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -6 gpurun_out/r2_pytest_gpu.log | cut -c1-400
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
