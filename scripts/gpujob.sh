cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_n2.log 2>&1; tail -8 gpurun_out/pytest_n2.log)
(BENCH_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_c3_n2.json 2> gpurun_out/bench_c3_n2.err; tail -12 gpurun_out/bench_c3_n2.err; head -c 5000 gpurun_out/bench_c3_n2.json)
