cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_full.log 2>&1; tail -6 gpurun_out/pytest_full.log)
for w in 1 2 4 8; do timeout 300 python scripts/rank_share.py $w 0 c3; done 2>&1 | grep "^{" | tee gpurun_out/rank_share.jsonl
(ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2_share8.csv -s 140 -c 60 env REPS=6 python scripts/rank_share.py 8 0 c3 > gpurun_out/ncu_share8.log 2>&1; tail -2 gpurun_out/ncu_share8.log)
