# single-GPU: ncu of the dominant kernel (probe and inside bench.py), launch list of a bench step, sanitizer passes
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:pair_sym -s 2 -c 1 -f -o gpurun_out/prof_r2_sym python scripts/probe_sym.py 96000 > gpurun_out/ncu_r2_sym.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pair_sym -s 24 -c 1 -f -o gpurun_out/prof_r2_bench_sym python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_r2_bench_sym.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2_bench.csv -s 500 -c 120 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_r2_launches.log 2>&1
ncu --set full --clock-control none -k regex:fiber_gemv -s 12 -c 2 -f -o gpurun_out/prof_r2_fiber python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_r2_fiber.log 2>&1
(compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_fiberops.py -m gpu -x -q -k "ragged or reference_shaped or long_fiber" 2>&1 | tail -4) > gpurun_out/san_race_fiber.txt
(compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_symmetric.py tests/test_gpu_exclusion.py -m gpu -x -q -k "2048-300 or 3000-1 or duplicate or 3000-1-1 or 700-333" 2>&1 | tail -4) > gpurun_out/san_race_sym.txt
(compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_mflow.py tests/test_gpu_fiberops.py tests/test_gpu_exclusion.py -m gpu -x -q -k "40-nodes0 or velocity_at_targets or reference_shaped or ragged or 700-333 or 3000-1-1" 2>&1 | tail -4) > gpurun_out/san_mem.txt
tail -3 gpurun_out/san_*.txt; ls -la gpurun_out/*.ncu-rep
