#!/usr/bin/env python
"""SASS-exact constants of the lane-instruction model ("replace W/D/R by exact
counts from the SASS") from an `ncu --set full --import-source on` capture of play_kernel.

    python scripts/sass_model.py REPORT.ncu-rep LANE_ROLLS RNG_WORDS [NAME]  ->  JSON on stdout

(LANE_ROLLS and RNG_WORDS are totals[3] and totals[5] of the captured launch: rolls played and
64-bit outputs the reference's generators would have produced.)  Executed thread instructions are
attributed by the CUDA source line the report correlates them with:
  W  per PCG64-DXSM word: everything the face-queue top-up executes (the G block of play_kernel and
     the rng.cuh lines inlined into it: output function, LCG step, the two Lemire products per word,
     packing, the queue insert), divided by the words it COMPUTED.  The top-up computes a few per
     cent more words than the reference's generators would have produced (a lane is topped up three
     words at a time and a game ends with a few codes unread): that surplus is execution overhead,
     not algorithmic work
  D  0: a die costs nothing beyond its share of a word (two dice per word) and of the roll
  R  per roll: everything else (taking the roll's dice off the queue, score lookup, discards,
     counters, hot dice / entry gate / final round / keep decision, turn switch, seat staging, lane
     refill) at the lane occupancy the kernel achieves: R = (all thread instructions - top-up) / rolls
  E  executed thread instructions per roll of the capture (all of them)
so that W words + R rolls is the algorithmic work and E rolls what the kernel executes.
"""
from __future__ import annotations

import csv
import io
import json
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def main() -> None:
    rep, rolls, words = sys.argv[1], float(sys.argv[2]), float(sys.argv[3])
    name = sys.argv[4] if len(sys.argv) > 4 else "play_kernel"
    src = (ROOT / "farkle_ii_b200/csrc/play.cuh").read_text().splitlines()
    g_lo = next(i + 1 for i, l in enumerate(src) if "// ================= G:" in l)
    g_hi = next(i + 1 for i, l in enumerate(src) if "// ================= P:" in l)
    words_line = next(i + 1 for i, l in enumerate(src) if "fq_bits += FQ_GEN_BITS;" in l)  # one add per served lane
    helper_line = next(i + 1 for i, l in enumerate(src) if "mul.wide.u32 t, %2, 6" in l)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # The view lists a SASS instruction once per level of its inline stack (callee line and call
    # site): classify every ADDRESS once, by all the source lines it appears under.
    hdr, cur_file, cur_line = None, "?", 0
    tags: dict[str, set] = {}
    thread_inst: dict[str, int] = {}
    text: dict[str, str] = {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].rsplit("/", 1)[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            it, ia, isrc = hdr.index("Thread Instructions Executed"), hdr.index("Address"), hdr.index("Source")
            continue
        if hdr is None or not r or len(r) < len(hdr):
            continue
        extra = len(r) - len(hdr)      # source text with quotes and commas splits into extra fields
        if r[0].isdigit():
            cur_line = int(r[0])
            continue
        addr = r[ia + extra] if not r[ia].startswith("0x") else r[ia]
        if not addr.startswith("0x"):
            continue
        kind = None
        # the top-up: its own lines, the generator functions inlined into it (rng.cuh; their few other
        # uses, the replay paths of a rejected half, execute a handful of times per launch) and the
        # Lemire product helper
        if (cur_file == "play.cuh" and (g_lo <= cur_line < g_hi or cur_line == helper_line)) or cur_file == "rng.cuh":
            kind = "G"
            if cur_file == "play.cuh" and cur_line == words_line:
                tags.setdefault(addr, set()).add("WORDS")
        tags.setdefault(addr, set()).add(kind)
        thread_inst[addr] = int(r[it + extra] or 0)
        text[addr] = r[isrc + extra] if isrc + extra < len(r) else ""
    topup = sum(t for a_, t in thread_inst.items() if "G" in tags[a_])
    # lane-firings: the add that lengthens the queue executes once per served lane
    firing = [t for a_, t in thread_inst.items() if "WORDS" in tags[a_]]
    computed = 3.0 * sum(firing)  # (the loop body is instantiated twice: one add per copy)
    total = sum(thread_inst.values())
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    m = dict(zip(rr[0], rr[2]))
    # thread instructions = warp instructions x average active threads per instruction
    executed = float(m["smsp__inst_executed.sum"]) * float(m["smsp__thread_inst_executed_per_inst_executed.ratio"])
    print(json.dumps({"kernel": name, "W": round(topup / max(computed, 1.0), 2), "D": 0.0,
                      "R": round((executed - topup) / rolls, 2), "E": round(executed / rolls, 2),
                      "words_computed_per_consumed": round(computed / words, 4),
                      "correlated_thread_inst_per_roll": round(total / rolls, 2),
                      "lane_rolls": rolls, "rng_words": words,
                      "issue_active_pct": float(m["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
                      "threads_per_inst": float(m["smsp__thread_inst_executed_per_inst_executed.ratio"]),
                      "source": rep.rsplit("/", 1)[-1]}))


if __name__ == "__main__":
    main()
