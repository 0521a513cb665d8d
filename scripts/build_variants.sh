#!/bin/bash
# Build tuning variants of libskelly_b200.so into gpurun_variants/<name>/ (git-ignored, travels to the GPU box).
# usage: build_variants.sh name:"-DFLAG=.. -DFLAG=.." ...
set -e
cd "$(dirname "$0")/.."
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  out="$PWD/gpurun_variants/$name"
  mkdir -p "$out"
  make -s -C skellysim_b200/csrc OUT="$out" EXTRA="$flags" > "$out/build.log" 2>&1 || { tail -5 "$out/build.log"; exit 1; }
  regs=$(grep -A3 "pair_sym_kernel" "$out/skb_runtime.ptxas.log" | grep -o "Used [0-9]* registers" | head -1)
  spill=$(grep -A2 "pair_sym_kernel" "$out/skb_runtime.ptxas.log" | grep -o "[0-9]* bytes spill stores" | head -1)
  echo "$name [$flags]: $regs, $spill"
  rm -f "$out"/*.o
done
