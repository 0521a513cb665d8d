cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python scripts/probe_small_targets.py > gpurun_out/probe_small.log 2>&1; tail -3 gpurun_out/probe_small.log
