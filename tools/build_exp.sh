#!/bin/bash
# build an experimental variant of the library: tools/build_exp.sh NAME [-DEXP_FLAG ...]
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p exp
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" \
  -o exp/lib_$name.so cnn-super-resolution_b200/csrc/srcnn.cu
