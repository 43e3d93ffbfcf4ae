cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -m pytest tests -q -m gpu -x -k "workspace or tiny_v2 or rmse" --timeout 150 2>&1 | tail -2
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
