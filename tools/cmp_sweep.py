"""Print two attention sweeps (tools/sweep_attn.py outputs) side by side: fraction of the HBM roofline, forward / backward."""
import json, sys
a = json.load(open(sys.argv[1]))["points"]
b = json.load(open(sys.argv[2]))["points"] if len(sys.argv) > 2 else a
print("ws T heads shift |  fwd A -> B   |  bwd A -> B")
for x, y in zip(a, b):
    print("%2d %d %5d %5d | %.3f -> %.3f | %.3f -> %.3f" % (x["ws"], x["T"], x["heads"], x["shift"], x.get("fwd_frac", 0), y.get("fwd_frac", 0),
                                                       x.get("bwd_frac", 0), y.get("bwd_frac", 0)))
fl = lambda p, k: min(q.get(k, 0) for q in p)
print("floor: fwd %.3f -> %.3f, bwd %.3f -> %.3f" % (fl(a, "fwd_frac"), fl(b, "fwd_frac"), fl(a, "bwd_frac"), fl(b, "bwd_frac")))
