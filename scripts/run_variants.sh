#!/bin/bash
# GPU box: time every variant under gpurun_variants/ with scripts/probe_sym.py -> gpurun_out/sym_variants.jsonl
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/sym_variants.jsonl
for d in gpurun_variants/*/; do
  lib="$PWD/${d}libskelly_b200.so"
  [ -f "$lib" ] || continue
  SKB_LIBRARY="$lib" timeout 300 python scripts/probe_sym.py "$@" 2>&1 | tail -1 | tee -a gpurun_out/sym_variants.jsonl
done
