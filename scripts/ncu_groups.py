#!/usr/bin/env python
"""Group an ncu source-page profile of play_kernel by phase of the loop body.

    python scripts/ncu_groups.py REPORT.ncu-rep KERNEL_SUBSTRING N_LANE_ROLLS
"""
import csv, io, re, subprocess, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
import ncu_lines as nl

rep, kernel, rolls = sys.argv[1], sys.argv[2], float(sys.argv[3])
src = (nl.ROOT / "farkle_ii_b200/csrc/play.cuh").read_text().splitlines()
marks = [(i + 1, m.group(1)) for i, l in enumerate(src)
         if (m := re.search(r"// =+ ([A-Z]): ", l))]
tail = next(i + 1 for i, l in enumerate(src) if "totals: warp shuffle" in l)
def phase(ln):
    if ln >= tail: return "Z totals"
    cur = "0 prologue"
    for start, name in marks:
        if ln >= start: cur = name
    return cur
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = next(r for r in rows if "Address" in r and "Source" in r)
ia, ii, it, ism = (hdr.index(c) for c in ("Address", "Instructions Executed", "Thread Instructions Executed", "# Samples"))
body = [r for r in rows if len(r) == len(hdr) and r[ia].startswith("0x")]
base = int(body[0][ia], 16)
lines = {off: (f, ln) for off, f, ln in nl.sass_lines(kernel)}
agg = {}
for r in body:
    f, ln = lines.get(int(r[ia], 16) - base, ("?", 0))
    g = phase(ln) if f == "play.cuh" else f
    a = agg.setdefault(g, [0, 0, 0]); a[0] += int(r[ii]); a[1] += int(r[it]); a[2] += int(r[ism])
ti = sum(a[0] for a in agg.values()); ts = sum(a[2] for a in agg.values())
print(f"total warp-inst {ti:,}; per warp-iteration (32 lane-rolls): {ti * 32 / rolls:.1f}")
for g, (i, t, s) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{g:28s} warp-inst/iter {i * 32 / rolls:7.1f} ({100 * i / ti:5.1f}%)  thread-inst/roll {t / rolls:6.1f}"
          f"  thr/inst {t / max(i, 1):5.1f}  samples {100 * s / ts:5.1f}%")
