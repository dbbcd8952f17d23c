#!/bin/bash
# A/B of decision-loop variants (tools/ab_build.sh) on one GPU: per-decision cost from the in-kernel trace + the c4 / c3 bench lines.
tag=${1:-ab}; shift
mkdir -p gpurun_out
for v in "$@"; do
  ISL_LIB=$PWD/tools/_ab/$v.so timeout 300 python tools/chain_cost.py > gpurun_out/${tag}_${v}_chain_cost.txt 2>&1; echo "$v chain_cost rc=$?"; tail -4 gpurun_out/${tag}_${v}_chain_cost.txt
done
