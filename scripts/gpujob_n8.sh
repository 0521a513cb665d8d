# 8-GPU box: the weak-scaled headline as the driver launches it, C4, N=4, and the multi-GPU tests
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
(BENCH_VERBOSE=1 timeout 600 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_c3_n8.json 2> gpurun_out/bench_c3_n8.err; tail -3 gpurun_out/bench_c3_n8.err; head -c 1500 gpurun_out/bench_c3_n8.json; echo)
(BENCH_VERBOSE=1 timeout 600 $TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 3 --no-inprocess > gpurun_out/bench_c3_n4.json 2> gpurun_out/bench_c3_n4.err; tail -2 gpurun_out/bench_c3_n4.err; head -c 600 gpurun_out/bench_c3_n4.json; echo)
(BENCH_VERBOSE=1 timeout 900 $TR --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --workload c4 --steps 5 --warmup 3 --no-extras > gpurun_out/bench_c4_n8.json 2> gpurun_out/bench_c4_n8.err; tail -3 gpurun_out/bench_c4_n8.err; head -c 1200 gpurun_out/bench_c4_n8.json; echo)
(timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_dense.py tests/test_gpu_mflow.py -m gpu -x -q > gpurun_out/pytest_n8.log 2>&1; tail -6 gpurun_out/pytest_n8.log)
(BENCH_VERBOSE=1 timeout 600 $TR --nproc-per-node 8 --master-port 29524 bench.py --gpus 8 --steps 20 --warmup 3 --exchange nccl --no-extras > gpurun_out/bench_c3_n8_nccl.json 2> gpurun_out/bench_c3_n8_nccl.err; tail -2 gpurun_out/bench_c3_n8_nccl.err; head -c 600 gpurun_out/bench_c3_n8_nccl.json; echo)
