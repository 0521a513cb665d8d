cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_full.log 2>&1; tail -6 gpurun_out/pytest_full.log)
for w in 2 4 8; do timeout 300 python scripts/rank_share.py $w 0 c3; done 2>&1 | grep "^{" | tee gpurun_out/rank_share.jsonl
python scripts/probe_sym.py 96000 32000 2>&1 | tail -1
(BENCH_VERBOSE=1 timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c3_n1.json 2> gpurun_out/bench_c3_n1.err; tail -3 gpurun_out/bench_c3_n1.err; head -c 700 gpurun_out/bench_c3_n1.json; echo)
(timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_n1.json 2>/dev/null; head -c 400 gpurun_out/bench_ref_n1.json; echo)
