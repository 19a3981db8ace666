"""Stall-reason breakdown of an `ncu --set full --import-source on` capture per block of SASS instructions (a warp-specialised
kernel: each role's code is one contiguous range):
    ncu -i X.ncu-rep --page source --csv --print-source sass | python tools/ncu_roles.py [block]"""
import collections, csv, sys
blk = int(sys.argv[1]) if len(sys.argv) > 1 else 250
rows = list(csv.reader(l for l in sys.stdin if l.startswith('"')))
hdr = next(r for r in rows if r and r[0] == "Address")
body = rows[rows.index(hdr) + 1:]
isrc, ismp = hdr.index("Source"), hdr.index("# Samples")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ismp] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
for a in range(0, len(body), blk):
    seg = body[a:a + blk]
    n = sum(int(r[ismp] or 0) for r in seg)
    rs = collections.Counter()
    for r in seg:
        for i in stall:
            rs[hdr[i][6:]] += int(r[i] or 0)
    ops = collections.Counter((r[isrc].split()[1] if r[isrc].startswith("@") else r[isrc].split()[0]).split(".")[0] for r in seg if r[isrc])
    print(f"{a:6d} {100.0 * n / max(tot, 1):5.1f}%  " + ", ".join(f"{k} {v}" for k, v in rs.most_common(5) if v) + "   | " +
          " ".join(f"{k}:{v}" for k, v in ops.most_common(4)))
