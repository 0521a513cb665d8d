cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_fiberops.py tests/test_gpu_mflow.py tests/test_gpu_solve.py tests/test_gpu_flow.py -m gpu -x -q > gpurun_out/pytest_overlap.log 2>&1; tail -8 gpurun_out/pytest_overlap.log)
for ov in 1 0; do
  SKB_OVERLAP=$ov timeout 300 python bench.py --steps 5 --warmup 3 --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(json.dumps({'overlap': $ov, 'ms_per_matvec': d['ms_per_matvec'], 'e2e_ms': d['e2e'].get('ms_per_matvec'), 'sym_ms': d['roofline'].get('kernel_ms'), 'err': d['accuracy']['max_rel_err_vs_oracle'], 'launches': d.get('gpu_launches')}))" | tee -a gpurun_out/overlap.jsonl
done
