"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel total time / launches / share of the
LAST training step in the file, plus the family shares bench.py reports (roofline.family_time_shares).
    python tools/summarize_launches.py gpurun_out/launches.csv [steps_in_file] > profiles/..._summary.txt"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        rows.append((r["Kernel Name"], us))
# the last step = the launches after the last-but-one optimizer kernel group; approximate by the final 1/steps of the stswin launches
ours = [i for i, (k, _) in enumerate(rows) if "stswin" in k]
per_step = len(ours) // steps
start = ours[len(ours) - per_step]
# include the at:: kernels between our launches of that step and the optimizer kernels after them
sel = rows[start:]


def family(k):
    if "gemm_kernel" in k:
        return "gemm_wgrad" if re.search(r"gemm_kernel<\d, \d, 4>", k) else "gemm"
    for name, fam in (("winattn_bwd", "winattn_bwd"), ("winattn_fwd", "winattn_fwd"), ("ln_bwd", "layernorm_bwd"), ("ln_fwd", "layernorm_fwd"),
                      ("transpose", "transpose"), ("copy_strided", "copy")):
        if name in k:
            return fam
    return None


tot = sum(us for _, us in sel)
agg = defaultdict(lambda: [0.0, 0])
fam = defaultdict(float)
for k, us in sel:
    short = re.sub(r"\(.*", "", k)[:100]
    agg[short][0] += us
    agg[short][1] += 1
    f_ = family(k)
    if f_:
        fam[f_] += us
st = sum(fam.values())
print(f"# {path}: last of {steps} training steps, per kernel: total us, launches, share (cold-cache, serialised times under ncu:")
print("# compare SHARES with bench.py's roofline.family_time_shares, not absolute times)")
print(f"# total {tot:.0f} us over {len(sel)} launches; stswin kernels {st:.0f} us = {100 * st / tot:.1f}%")
print("# family shares of the stswin kernels: " + ", ".join(f"{k} {v / st:.4f}" for k, v in sorted(fam.items(), key=lambda kv: -kv[1])))
for k, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{us:10.1f} us {n:5d} {100 * us / tot:5.1f}%  {k}")
