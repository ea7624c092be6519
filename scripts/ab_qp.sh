#!/bin/bash
# A/B of QP-kernel builds on one GPU: scripts/ab_qp.sh <tag> lib1.so lib2.so ...   (libraries under racing-lmpc-ros2_b200/csrc/)
TAG=$1; shift
mkdir -p gpurun_out
for lib in "$@"; do
  for rep in 1 2; do
    echo "== $lib (rep $rep)" >> gpurun_out/ab_${TAG}.txt
    LMPC_B200_LIB=$PWD/racing-lmpc-ros2_b200/csrc/$lib timeout 300 python scripts/time_qp.py 1024 8192 >> gpurun_out/ab_${TAG}.txt 2>&1
  done
done
cat gpurun_out/ab_${TAG}.txt
