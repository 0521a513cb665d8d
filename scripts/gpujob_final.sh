cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_full.log 2>&1; tail -5 gpurun_out/pytest_full.log)
(timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log)
(timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; head -c 400 gpurun_out/bench_ref_n1.json; echo)
(timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-extras > gpurun_out/bench_c2_n1.json 2> gpurun_out/bench_c2_n1.err; head -c 400 gpurun_out/bench_c2_n1.json; echo)
timeout 900 bash scripts/run_variants.sh
