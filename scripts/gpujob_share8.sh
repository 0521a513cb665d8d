cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for w in 8 4 2; do timeout 200 python scripts/rank_share.py $w 0 c3 2>&1 | tail -1 | tee -a gpurun_out/rank_share_new.jsonl; done
SKB_OVERLAP=0 timeout 200 python scripts/rank_share.py 8 0 c3 2>&1 | tail -1 | tee -a gpurun_out/rank_share_new.jsonl
REPS=3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_share8_new.csv -s 150 -c 60 python scripts/rank_share.py 8 0 c3 > gpurun_out/ncu_share8.log 2>&1
tail -2 gpurun_out/ncu_share8.log
