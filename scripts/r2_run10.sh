cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python scripts/diag_v1.py 2>&1 | grep -v Warn | tail -10
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -8 gpurun_out/r2_pytest_gpu.log | cut -c1-300
