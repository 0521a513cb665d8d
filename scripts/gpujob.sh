cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_flow.py tests/test_gpu_mflow.py tests/test_gpu_fiberops.py tests/test_gpu_solve.py -m gpu -x -q > gpurun_out/pytest_cross.log 2>&1; tail -8 gpurun_out/pytest_cross.log)
(SKB_CROSS=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('cross off', d['ms_per_matvec'], d['accuracy']['max_rel_err_vs_oracle'], d.get('error'))")
(timeout 600 python bench.py --steps 5 --warmup 3 --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('cross auto', d['ms_per_matvec'], d['accuracy']['max_rel_err_vs_oracle'], d.get('error'), d['launches_per_matvec_per_rank'])")
for w in 2 4 8; do timeout 300 python scripts/rank_share.py $w 0 c3; done 2>&1 | grep "^{" | tee gpurun_out/rank_share_cross.jsonl
