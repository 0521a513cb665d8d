cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_fiberops.py tests/test_gpu_solve.py tests/test_gpu_mflow.py -m gpu -x -q > gpurun_out/pytest_fiber2.log 2>&1; tail -3 gpurun_out/pytest_fiber2.log)
timeout 200 python scripts/probe_fiber_ops.py 2>&1 | tail -1 | tee gpurun_out/fiber_ops_probe2.json
(timeout 500 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_fiberops.py -m gpu -x -q -k "ragged or long_fiber" > gpurun_out/racecheck_fiber2.log 2>&1; tail -4 gpurun_out/racecheck_fiber2.log)
