"""Top stall locations of an `ncu --set full --import-source on` capture, by SASS instruction:
    ncu -i X.ncu-rep --page source --csv --print-source sass | python tools/ncu_hot_sass.py [top_n]"""
import csv
import sys

top = int(sys.argv[1]) if len(sys.argv) > 1 else 25
rows = list(csv.reader(l for l in sys.stdin if l.startswith('"')))
hdr = next(r for r in rows if r and r[0] == "Address")
body = rows[rows.index(hdr) + 1:]
ia, isrc, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
total = sum(int(r[ismp] or 0) for r in body)
order = sorted(range(len(body)), key=lambda k: -int(body[k][ismp] or 0))[:top]
print(f"total samples {total}")
for k in sorted(order):
    r = body[k]
    n = int(r[ismp] or 0)
    reasons = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    print(f"{k:5d} {100.0 * n / max(total, 1):5.1f}%  {r[isrc][:90]:90s} {reasons}")
