# N-GPU box: the weak-scaled headline as the driver launches it (+ C4 at N = 8)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
(BENCH_VERBOSE=1 timeout 500 $TR --nproc-per-node $N --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_c3_n$N.json 2> gpurun_out/bench_c3_n$N.err; tail -2 gpurun_out/bench_c3_n$N.err; grep "^{" gpurun_out/bench_c3_n$N.json | head -c 700; echo)
if [ "$N" = "8" ]; then
(BENCH_VERBOSE=1 timeout 500 $TR --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --workload c4 --steps 5 --warmup 3 --no-extras > gpurun_out/bench_c4_n8.json 2> gpurun_out/bench_c4_n8.err; tail -2 gpurun_out/bench_c4_n8.err; grep "^{" gpurun_out/bench_c4_n8.json | head -c 500; echo)
fi
if [ "$N" = "2" ]; then
(timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_dense.py tests/test_gpu_mflow.py -m gpu -x -q > gpurun_out/pytest_n2.log 2>&1; tail -4 gpurun_out/pytest_n2.log)
fi
