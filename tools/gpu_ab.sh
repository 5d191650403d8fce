#!/bin/bash
# A/B of library variants built with tools/build_exp.sh: tools/gpu_ab.sh TAG name1 name2 ... [-- pytest -k expr]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=$1; shift
for n in "$@"; do
  [ "$n" = "--" ] && break
  SRCNN_B200_LIB=exp/lib_$n.so timeout 300 python tools/ab_infer.py 30 2>&1 | tail -1 | tee -a gpurun_out/${T}_ab.txt
  SRCNN_B200_LIB=exp/lib_$n.so timeout 300 python tools/train_time.py 2>&1 | tail -3 | tee -a gpurun_out/${T}_ab.txt
done
