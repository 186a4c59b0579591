#!/usr/bin/env python
"""SASS-exact constants of the lane-instruction model ("replace W/D/R by exact
counts from the SASS") from an `ncu --set full --import-source on` capture of play_kernel.

    python scripts/sass_model.py REPORT.ncu-rep LANE_ROLLS [NAME]  ->  JSON on stdout

Executed thread instructions are attributed by the CUDA source line the report correlates them with:
  W  per PCG64-DXSM word COMPUTED (three per roll): the output function and the LCG step in rng.cuh,
     the step calls and the four state selects per word in play.cuh
  D  per die SLOT (six per roll): the FB_DIE lines and the min-reduction of the Lemire leftovers
  R  per roll: everything else the kernel executes (score lookup, discards, counters, hot dice /
     entry gate / final round / keep decision, turn switch, seat staging, lane refill), i.e.
     R = (all thread instructions - 3 W rolls - 6 D rolls) / rolls
so that 3 W + 6 D + R is the executed thread-instruction count per roll of the capture, and
W words + D dice + R rolls (words and dice actually CONSUMED, counted by the kernel) is the
algorithmic work: what remains when the unused third word and the unused die slots are not charged.
"""
from __future__ import annotations

import csv
import io
import json
import re
import subprocess
import sys


def main() -> None:
    rep, rolls = sys.argv[1], float(sys.argv[2])
    name = sys.argv[3] if len(sys.argv) > 3 else "play_kernel"
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # The view lists a SASS instruction once per level of its inline stack (callee line and call
    # site): classify every ADDRESS once, by all the source lines it appears under.
    hdr, cur_file, cur_text = None, "?", ""
    tags: dict[str, set] = {}
    thread_inst: dict[str, int] = {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].rsplit("/", 1)[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            it, ia = hdr.index("Thread Instructions Executed"), hdr.index("Address")
            continue
        if hdr is None or not r or len(r) < len(hdr):
            continue
        extra = len(r) - len(hdr)      # source text with quotes and commas splits into extra fields
        if r[0].isdigit():
            cur_text = ",".join(r[1:2 + extra])
            continue
        addr = r[ia + extra] if not r[ia].startswith("0x") else r[ia]
        if not addr.startswith("0x"):
            continue
        kind = None
        if cur_file == "rng.cuh":
            kind = "W"
        elif cur_file == "play.cuh" and re.search(r"shi = c[123]|slo = c[123]", cur_text):
            kind = "W"
        elif cur_file == "play.cuh" and "FB_DIE(" in cur_text:
            kind = "D"
        elif cur_file == "math_functions.hpp" and "umin" in cur_text:
            kind = "D"
        tags.setdefault(addr, set()).add(kind)
        thread_inst[addr] = int(r[it + extra] or 0)
    w = sum(t for a_, t in thread_inst.items() if "W" in tags[a_])
    d = sum(t for a_, t in thread_inst.items() if "D" in tags[a_] and "W" not in tags[a_])
    total = sum(thread_inst.values())
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    m = dict(zip(rr[0], rr[2]))
    # thread instructions = warp instructions x average active threads per instruction
    executed = float(m["smsp__inst_executed.sum"]) * float(m["smsp__thread_inst_executed_per_inst_executed.ratio"])
    W, D = w / (3 * rolls), d / (6 * rolls)
    R = (executed - w - d) / rolls
    print(json.dumps({"kernel": name, "W": round(W, 2), "D": round(D, 2), "R": round(R, 2),
                      "executed_thread_inst_per_roll": round(executed / rolls, 2),
                      "correlated_thread_inst_per_roll": round(total / rolls, 2),
                      "lane_rolls": rolls, "issue_active_pct": float(m["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
                      "threads_per_inst": float(m["smsp__thread_inst_executed_per_inst_executed.ratio"]),
                      "source": rep.rsplit("/", 1)[-1]}))


if __name__ == "__main__":
    main()
