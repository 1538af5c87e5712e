#!/bin/bash
# Tuning builds of the library: tools/variants/<name>/libmouse_b200.so, one per line of VARIANTS ("name flags...").
# (git-ignored; they travel to the GPU box with gpurun.)  Usage: tools/build_variants.sh [file-with-variants]
set -e
cd "$(dirname "$0")/.."
LIST=${1:-tools/variants.txt}
while read -r name flags; do
  [ -z "$name" ] && continue
  case "$name" in \#*) continue;; esac
  d=tools/variants/$name
  mkdir -p "$d"
  make -s -C mouse_b200/csrc BUILD="../../$d/build" OUT="../../$d/libmouse_b200.so" EXTRA="$flags" &
done < "$LIST"
wait
for f in tools/variants/*/build/p2_sweep.ptxas.log; do
  echo "== $f"; grep -A2 "p2_sweep_half_kernelILi[468]ELb0" "$f" | grep "Used\|spill" | tr '\n' ' '; echo
done
