#!/usr/bin/env python
"""Per-source-line view of an ncu report (no GUI): joins `ncu --page source --csv` (SASS rows,
in program order) with `nvdisasm -g` line markers of the same cubin.

    python scripts/ncu_lines.py REPORT.ncu-rep KERNEL_SUBSTRING [TOP]

Prints the hottest source lines by warp instructions executed, with average active threads
per instruction and stall samples.  Inlined code is attributed to the innermost line.
"""
from __future__ import annotations

import csv
import io
import re
import subprocess
import sys
import tempfile
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def sass_lines(kernel: str) -> list[tuple[int, str, int]]:
    """[(offset, file, line)] for the first kernel whose mangled name contains `kernel`."""
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", str(ROOT / "farkle_ii_b200/libfarkle_b200.so")],
                       cwd=td, check=True, capture_output=True)
        cubin = next(Path(td).glob("*.cubin"))
        text = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True,
                              text=True).stdout
    out, active, cur = [], False, ("?", 0)
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
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*)", ln)
        if m:
            out.append((int(m.group(1), 16), cur[0], cur[1]))
    return out


def main() -> None:
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next(r for r in rows if "Address" in r and "Source" in r)
    ia, ii, it, ism = (hdr.index(c) for c in ("Address", "Instructions Executed",
                                              "Thread Instructions Executed", "# Samples"))
    ipred = hdr.index("Predicated-On Thread Instructions Executed")
    body = [r for r in rows if len(r) == len(hdr) and r[ia].startswith("0x")]
    base = int(body[0][ia], 16)
    lines = {off: (f, ln) for off, f, ln in sass_lines(kernel)}
    agg = defaultdict(lambda: [0, 0, 0, 0])
    for r in body:
        key = lines.get(int(r[ia], 16) - base, ("?", 0))
        a = agg[key]
        a[0] += int(r[ii]); a[1] += int(r[it]); a[2] += int(r[ism]); a[3] += int(r[ipred])
    tot_i = sum(a[0] for a in agg.values())
    tot_t = sum(a[1] for a in agg.values())
    tot_s = sum(a[2] for a in agg.values())
    print(f"warp-inst {tot_i:,}  thread-inst {tot_t:,}  avg threads/inst {tot_t / tot_i:.2f}  "
          f"samples {tot_s:,}")
    src_cache: dict[str, list[str]] = {}
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        if f not in src_cache:
            p = ROOT / "farkle_ii_b200/csrc" / f
            src_cache[f] = p.read_text().splitlines() if p.exists() else []
        text = src_cache[f][ln - 1].strip() if 0 < ln <= len(src_cache[f]) else ""
        print(f"{f:12s}:{ln:4d} inst {100 * a[0] / tot_i:5.2f}%  thr/inst {a[1] / max(a[0], 1):5.1f}"
              f"  pred-on {a[3] / max(a[0], 1):5.1f}  samples {100 * a[2] / max(tot_s, 1):5.2f}% | {text[:90]}")


if __name__ == "__main__":
    main()
