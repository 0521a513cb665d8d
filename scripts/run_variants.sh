#!/bin/bash
# GPU box: time every variant under gpurun_variants/ -> gpurun_out/variants.jsonl (symmetric-kernel probe + a short bench)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/variants.jsonl
for d in gpurun_variants/*/; do
  lib="$PWD/${d}libskelly_b200.so"
  [ -f "$lib" ] || continue
  SKB_LIBRARY="$lib" timeout 300 python scripts/probe_sym.py 96000 2>&1 | tail -1 | tee -a gpurun_out/variants.jsonl
  [ -n "$PROBE_ONLY" ] || SKB_LIBRARY="$lib" timeout 300 python bench.py --steps 5 --warmup 3 --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(json.dumps({'lib': '$d', 'ms_per_matvec': d['ms_per_matvec'], 'err': d['accuracy']['max_rel_err_vs_oracle']}))" | tee -a gpurun_out/variants.jsonl
done
