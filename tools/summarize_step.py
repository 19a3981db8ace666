"""One training step out of an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py --no-graph: the launches
between two adam_tick kernels (the step counter of FusedAdam), per kernel and per family.
    python tools/summarize_step.py launches.csv [which_step] > profiles/..._launch_list_summary.txt"""
import collections, csv, re, sys
path = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 1          # step after the which-th tick
lines = [l for l in open(path, errors="ignore") if l.startswith('"')]
rows = []
for r in csv.DictReader(lines):
    if r["Metric Name"] == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
        rows.append((r["Kernel Name"], v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)))
ticks = [i for i, (k, _) in enumerate(rows) if "adam_tick" in k]
sel = rows[ticks[which - 1] + 1:ticks[which] + 1]
tot = sum(u for _, u in sel)
agg = collections.defaultdict(lambda: [0, 0.0])
for k, u in sel:
    k = re.sub(r"\(.*", "", k).replace("<unnamed>::", "")
    if "stswin" not in k:
        k = "ATen: " + k[:70]
    agg[k][0] += 1; agg[k][1] += u
print("ncu launch list of ONE eager training step of the default bench (configs[1], 8 clips; ncu --metrics gpu__time_duration.sum")
print("--clock-control none: per-launch times are cold-cache and serialised -- compare SHARES).  python bench.py --steps 1 --warmup 1")
print("--no-graph; the launches between two adam_tick kernels")
print(f"{len(sel)} launches, {tot / 1e3:.2f} ms summed\n")
for k, (n, u) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:100s} {n:4d} launches {u:10.1f} us {100 * u / tot:6.2f} %")
ours = sum(u for k, (n, u) in agg.items() if "stswin" in k)
print(f"\nstswin:: kernels {100 * ours / tot:.2f} % of the step, everything else (ATen glue: the synthetic loss, fills, dtype casts) {100 - 100 * ours / tot:.2f} %")
fams = (("gemm_kernel", None), ("winattn_bwd", "winattn_bwd"), ("winattn_fwd", "winattn_fwd"), ("ln_bwd", "layernorm_bwd"), ("ln_fwd", "layernorm_fwd"),
        ("transpose", "transpose"), ("copy_strided", "copy"), ("adam", "adam"))
fs = collections.defaultdict(float)
for k, (n, u) in agg.items():
    for f, t in fams:
        if f in k:
            fs[t or ("gemm_wgrad" if re.search(r"gemm_kernel<\d, \d, 4>", k) else "gemm")] += u
            break
print("family shares of the stswin kernels under ncu: " + ", ".join(f"{k} {v / ours:.4f}" for k, v in sorted(fs.items(), key=lambda kv: -kv[1])))
