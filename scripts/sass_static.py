#!/usr/bin/env python
"""Static SASS instruction count per source-line range of a kernel in libfarkle_b200.so.

    python scripts/sass_static.py KERNEL_SUBSTRING [file:lo-hi ...]
"""
import re, subprocess, sys, tempfile
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
kernel = sys.argv[1]
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", str(ROOT / "farkle_ii_b200/libfarkle_b200.so")],
                   cwd=td, check=True, capture_output=True)
    cubin = next(Path(td).glob("*.cubin"))
    text = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True).stdout
active, cur = False, ("?", 0)
per_line, ops = Counter(), {}
for ln in text.splitlines():
    if ln.startswith(".text."):
        active = kernel in ln
        continue
    if not active:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (Path(m.group(1)).name, int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", ln)
    if m:
        per_line[cur] += 1
        ops.setdefault(cur, Counter())[m.group(2).split(".")[0]] += 1
print("total", sum(per_line.values()))
for spec in sys.argv[2:]:
    f, rng = spec.split(":")
    lo, hi = (int(x) for x in rng.split("-"))
    tot, mix = 0, Counter()
    for (ff, l), c in per_line.items():
        if ff == f and lo <= l <= hi:
            tot += c
            mix.update(ops[(ff, l)])
    print(f"{spec:28s} {tot:5d}  {dict(mix.most_common(8))}")
