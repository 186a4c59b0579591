#!/usr/bin/env python
"""Dynamic opcode histogram of a kernel from an ncu report (SASS rows of the source page):
warp instructions executed per opcode, per warp-iteration when LANE_ROLLS is given.

    python scripts/ncu_opcodes.py REPORT.ncu-rep [LANE_ROLLS]
"""
import csv, io, re, subprocess, sys
from collections import defaultdict

rep = sys.argv[1]
rolls = float(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if "Address" in r and "Source" in r)
ia, isrc, ii, it = (hdr.index(c) for c in ("Address", "Source", "Instructions Executed", "Thread Instructions Executed"))
agg = defaultdict(lambda: [0, 0])
for r in rows:
    if len(r) != len(hdr) or not r[ia].startswith("0x"):
        continue
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
    op = m.group(1) if m else "?"
    a = agg[op]
    a[0] += int(r[ii] or 0); a[1] += int(r[it] or 0)
tot = sum(a[0] for a in agg.values())
scale = 32 / rolls if rolls else 0
print(f"total warp-inst {tot:,}" + (f"; per warp-iteration {tot * scale:.1f}" if rolls else ""))
for op, (i, t) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"{op:28s} {100 * i / tot:6.2f}%  " + (f"{i * scale:7.2f}/iter  " if rolls else "") + f"thr/inst {t / max(i, 1):5.1f}")
