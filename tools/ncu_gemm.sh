set -x
M="dram__bytes_read.sum|dram__bytes_write.sum|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|gpu__time_duration.sum|l1tex__m_xbar2l1tex_read_bytes.sum |lts__t_sector_hit_rate.pct|sm__inst_executed.avg.per_cycle_elapsed|sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed|smsp__issue_active.avg.pct_of_peak_sustained_active|launch__registers_per_thread|sm__cycles_elapsed.avg.per_second|sm__throughput.avg.pct_of_peak_sustained_elapsed|lts__throughput.avg.pct_of_peak_sustained_elapsed"
for mode in 0 2; do
  ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 4 -c 1 -o gpurun_out/gemm_tma_m${mode}_r1w -f python tools/prof_gemm.py 163840 2048 512 $mode > /dev/null 2>&1
  ncu -i gpurun_out/gemm_tma_m${mode}_r1w.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys,re
rows=list(csv.reader(sys.stdin))
hdr,units,vals=rows[0],rows[1],rows[2]
pat=re.compile(r'$M'.replace(' ',''))
for h,u,v in zip(hdr,units,vals):
    if pat.fullmatch(h) or h in ('$M'.replace(' ','').split('|')): print(f'{h} [{u}] = {v}')
" > gpurun_out/gemm_tma_m${mode}_r1w_summary.txt
  cat gpurun_out/gemm_tma_m${mode}_r1w_summary.txt
done
for shape in "163840 2048 512 0" "163840 2048 512 2" "163840 2048 512 3 1" "163840 512 512 0" "163840 512 2048 1" "163840 1536 512 0" "8192 8192 8192 0"; do python tools/prof_gemm.py $shape; done 2>&1 | grep gemm
