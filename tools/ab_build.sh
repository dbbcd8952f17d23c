#!/bin/bash
# Builds A/B variants of libislplace.so under tools/_ab/ (engine.py: ISL_LIB=<path> picks one).  usage: tools/ab_build.sh name "-DFLAG=.. ..."
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/_ab
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared $2 -o tools/_ab/$1.so instaslice_b200/csrc/islplace.cu
echo built tools/_ab/$1.so
