#!/bin/sh
# builds the stand-alone micro-benchmarks next to their sources (binaries are git-ignored)
set -e
cd "$(dirname "$0")"
for f in *.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -lineinfo \
       "$f" ../../stswincl_b200/csrc/host_util.cu -o "$(basename "$f" .cu).bin"
done
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -Xcompiler -fPIC -shared tools/micro/spin.cu -o tools/micro/libspin.so   # tools/exp_coresident.py
