"""Instruction mix of one launch of an .ncu-rep (ncu --set full --import-source on): executed warp instructions per opcode and
per region of the SASS, from the source page.   python tools/ncu_opmix.py rep.ncu-rep [launch] [top]"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
launch = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(launch), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r]
hdr = rows[hi[0]]
isrc, iex, ithr, iadr, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("Address"), hdr.index("# Samples")
end = hi[1] - 1 if len(hi) > 1 else len(rows)
ops, thr, smp = Counter(), Counter(), Counter()
seen = set()
lines = []
for r in rows[hi[0] + 1:end]:
    if len(r) <= iex or not r[iex].isdigit() or r[iadr] in seen:
        continue
    seen.add(r[iadr])
    src = r[isrc].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    ops[op] += int(r[iex]); thr[op] += int(r[ithr]); smp[op] += int(r[ismp]) if r[ismp].isdigit() else 0
    lines.append((int(r[iex]), int(r[ithr]), src))
tot = sum(ops.values())
print(rows[0][1] if rows and len(rows[0]) > 1 else "", "launch", launch, "warp instructions", tot)
print("%-12s %14s %6s %8s %7s" % ("opcode", "warp inst", "%", "thr/inst", "smp %"))
ts = sum(smp.values()) or 1
for op, c in ops.most_common(top):
    print("%-12s %14d %6.2f %8.1f %7.2f" % (op, c, 100.0 * c / tot, thr[op] / max(c, 1), 100.0 * smp[op] / ts))
if "--popc" in sys.argv:
    print("POPC instructions by executed count:")
    for ex, th, src in sorted([l for l in lines if "POPC" in l[2]], reverse=True)[:40]:
        print("%12d  %5.1f  %s" % (ex, th / max(ex, 1), src))
