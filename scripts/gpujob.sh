cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:pair_sym -s 24 -c 1 -f -o gpurun_out/prof_r2_bench_sym python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_r2_bench_sym.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2_bench.csv -s 460 -c 110 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_r2_launches.log 2>&1
for w in 4 8; do timeout 300 python scripts/rank_share.py $w 0 c3; done 2>&1 | grep "^{" | tee gpurun_out/rank_share.jsonl
(ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2_share8.csv -s 130 -c 60 env REPS=6 python scripts/rank_share.py 8 0 c3 > gpurun_out/ncu_share8.log 2>&1)
timeout 900 python scripts/sweep_c5.py > gpurun_out/sweep_c5_r2.log 2>&1; tail -12 gpurun_out/sweep_c5_r2.log
