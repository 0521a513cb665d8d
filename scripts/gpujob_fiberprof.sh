cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python scripts/probe_fiber_ops.py 2>&1 | tail -1 | tee gpurun_out/fiber_ops_probe.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fiber_gemv -s 4 -c 2 -f -o gpurun_out/prof_r2_fiber_gemv python scripts/probe_fiber_ops.py 3 > gpurun_out/ncu_fiber.log 2>&1; tail -2 gpurun_out/ncu_fiber.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:driver -c 4 --csv --log-file gpurun_out/ref_gpu_kernel.csv python -c "
import bench, json
g = bench.make_system('c3', 1)
print(json.dumps(bench.ref_gpu_leg(g, reps=2)))" > gpurun_out/ref_gpu_kernel.log 2>&1; tail -3 gpurun_out/ref_gpu_kernel.csv | cut -c1-300
