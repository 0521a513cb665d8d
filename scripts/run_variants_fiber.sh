#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/fiber_variants.jsonl
for d in gpurun_variants/*/; do
  lib="$PWD/${d}libskelly_b200.so"
  [ -f "$lib" ] || continue
  SKB_LIBRARY="$lib" timeout 200 python scripts/probe_fiber_ops.py 30 2>&1 | tail -1 | tee -a gpurun_out/fiber_variants.jsonl
done
