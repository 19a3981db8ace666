#!/bin/sh
# builds the stand-alone micro-benchmarks next to their sources (binaries are git-ignored)
set -e
cd "$(dirname "$0")"
for f in *.cu; do
  [ "$f" = "spin.cu" ] && continue
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -lineinfo \
       "$f" ../../stswincl_b200/csrc/host_util.cu -o "$(basename "$f" .cu).bin"
done
# shared library for tools/exp_coresident.py
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -Xcompiler -fPIC -shared spin.cu -o libspin.so
