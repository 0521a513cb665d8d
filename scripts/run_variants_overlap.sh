#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/overlap_probe.jsonl
for d in gpurun_variants/*/; do
  lib="$PWD/${d}libskelly_b200.so"
  [ -f "$lib" ] || continue
  SKB_LIBRARY="$lib" timeout 300 python scripts/probe_overlap.py 2>&1 | tail -1 | tee -a gpurun_out/overlap_probe.jsonl
done
