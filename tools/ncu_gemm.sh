#!/bin/bash
# ncu --set full of one launch of the dense-layer kernel at the fc1 shape of a stage-1 block (bias and GELU modes),
# summaries into gpurun_out/, then event timings of the shapes of the bench step (no profiler attached).
for mode in 0 2; do
  ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 4 -c 1 -f \
      -o gpurun_out/gemm_tma_m${mode}_r1w python tools/prof_gemm.py 163840 2048 512 $mode > /dev/null 2>&1
  ncu -i gpurun_out/gemm_tma_m${mode}_r1w.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py \
      > gpurun_out/gemm_tma_m${mode}_r1w_summary.txt
  cat gpurun_out/gemm_tma_m${mode}_r1w_summary.txt
done
for shape in "163840 2048 512 0" "163840 2048 512 2" "163840 2048 512 3 1" "163840 512 512 0" "163840 512 2048 1" \
             "163840 1536 512 0" "8192 8192 8192 0"; do
  python tools/prof_gemm.py $shape
done 2>&1 | grep gemm
