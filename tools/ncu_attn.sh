#!/bin/bash
# ncu --set full of one launch of the final window-attention kernels (stage-1 and stage-2 bench geometry, 16 pair
# batches) and of one LayerNorm backward; summaries into gpurun_out/.
for stage in 1 2; do
  for k in fwd bwd; do
    ncu --set full --clock-control none --import-source on -k regex:winattn_${k} -s 3 -c 1 -f \
        -o gpurun_out/attn_${k}_s${stage}_r1z python tools/prof_attn.py $stage > /dev/null 2>&1
    echo "== winattn_${k} stage ${stage}"
    ncu -i gpurun_out/attn_${k}_s${stage}_r1z.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py \
        | tee gpurun_out/attn_${k}_s${stage}_r1z_summary.txt
  done
done
ncu --set full --clock-control none --import-source on -k regex:ln_bwd -s 3 -c 1 -f -o gpurun_out/ln_bwd_r1z python tools/prof_ln.py > /dev/null 2>&1
echo "== ln_bwd 163840 x 512 (+res +colsum)"
ncu -i gpurun_out/ln_bwd_r1z.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py | tee gpurun_out/ln_bwd_r1z_summary.txt
python tools/prof_attn.py 1; python tools/prof_attn.py 2
