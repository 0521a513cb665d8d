cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_full.log 2>&1; tail -5 gpurun_out/pytest_full.log)
(BENCH_VERBOSE=1 timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c3_n1.json 2> gpurun_out/bench_c3_n1.err; tail -2 gpurun_out/bench_c3_n1.err; head -c 500 gpurun_out/bench_c3_n1.json; echo)
ncu --set full --clock-control none --import-source on -k regex:pair_cross -s 3 -c 1 -f -o gpurun_out/prof_r2_cross python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_r2_cross.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2_bench.csv -s 500 -c 115 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_r2_launches.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
