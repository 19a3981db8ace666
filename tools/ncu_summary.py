"""Print the handful of `ncu --set full` metrics the profiles/ summaries quote, from an .ncu-rep:
    ncu -i X.ncu-rep --page raw --csv | python tools/ncu_summary.py"""
import csv
import sys

WANT = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__registers_per_thread",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second",
        "sm__inst_executed.avg.per_cycle_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
rows = list(csv.reader(l for l in sys.stdin if l.startswith('"')))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    for h, u, v in zip(hdr, units, vals):
        if h in WANT:
            print(f"{h} [{u}] = {v}")
