"""Per-kernel SASS evidence for profiles/sass_summary.txt: counts of the Blackwell tensor-core / TMEM / TMA instructions in
every kernel of libstswin_b200.so (cuobjdump -sass, sm_100a).
    python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "stswincl_b200", "libstswin_b200.so")
WANT = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "UBLKCP", "SYNCS", "HMMA", "FFMA", "MUFU"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
kern, counts, order = None, collections.defaultdict(collections.Counter), []
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        kern = re.sub(r"\(anonymous namespace\)::", "", kern)
        kern = re.sub(r"\(.*", "", kern)
        order.append(kern)
        continue
    if kern is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        counts[kern]["total"] += 1
        if op in WANT:
            counts[kern][op] += 1
print("SASS instruction counts per kernel of stswincl_b200/libstswin_b200.so (cuobjdump -sass, sm_100a)")
print("UTCHMMA = tcgen05.mma (bf16), LDTM/STTM = tcgen05.ld/st (TMEM), UTMALDG/UTMASTG/UTMAREDG = TMA load / store / reduce,")
print("UTCBAR = tcgen05.commit, UBLKCP = bulk copy, SYNCS = mbarrier ops\n")
cols = ["total"] + WANT
print("%-96s" % "kernel" + "".join("%9s" % c for c in cols))
tot = collections.Counter()
for k in order:
    c = counts[k]
    print("%-96s" % k[:95] + "".join("%9d" % c[x] for x in cols))
    tot.update(c)
print("%-96s" % "ALL" + "".join("%9d" % tot[x] for x in cols))
