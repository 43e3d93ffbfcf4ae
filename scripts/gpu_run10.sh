cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python scripts/timeline.py 1 gpurun_out/timeline_s1.json 2>&1 | tail -6
timeout 600 python scripts/timeline.py 2 gpurun_out/timeline_s2.json 2>&1 | tail -6
