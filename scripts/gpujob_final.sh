cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_full.log 2>&1; tail -5 gpurun_out/pytest_full.log)
(timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log)
(BENCH_VERBOSE=1 timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c3_n1.json 2> gpurun_out/bench_c3_n1.err; tail -2 gpurun_out/bench_c3_n1.err; head -c 600 gpurun_out/bench_c3_n1.json; echo)
