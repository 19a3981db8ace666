#!/bin/sh
# Debug build of the library with in-kernel timeline tracing (tools/trace_attn.py).  Not shipped.
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/_trace
for f in stswincl_b200/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -DSTSWIN_TRACE \
       -Xcompiler -fPIC -c "$f" -o tools/_trace/$(basename "$f" .cu).o &
done
wait
nvcc -shared -o tools/_trace/libstswin_trace.so tools/_trace/*.o
