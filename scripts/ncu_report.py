#!/usr/bin/env python
"""Summarise one `ncu --set full --import-source on` capture of play_kernel as markdown, from the
report alone (metrics, stall reasons, instructions per loop phase and per source line; the CUDA
source the report imported is what the lines are matched against, not the working tree).

    python scripts/ncu_report.py REPORT.ncu-rep "TITLE" LANE_ROLLS > profiles/rNN_play_kernel_kK.md

LANE_ROLLS = rolls/game x games of the profiled launch (scripts/profile_cell.py prints both).
"""
from __future__ import annotations

import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size",
    "launch__block_size", "launch__registers_per_thread",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__thread_inst_executed_pred_on_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]
STALLS = "smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio"


def ncu(*args: str) -> list[list[str]]:
    out = subprocess.run(["ncu", *args], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main() -> None:
    rep, title, rolls = sys.argv[1], sys.argv[2], float(sys.argv[3])
    raw = ncu("-i", rep, "--page", "raw", "--csv")
    m = dict(zip(raw[0], zip(raw[1], raw[2])))
    print(f"# {title}\n")
    print("Source: one `ncu --set full --clock-control none --import-source on` capture (scratch `.ncu-rep`, "
          "not tracked); summary by `scripts/ncu_report.py`.\n")
    print("| metric | value | unit |\n|---|---:|---|")
    for k in KEYS:
        if k in m:
            print(f"| `{k}` | {m[k][1]} | {m[k][0]} |")
    lanes = float(m["smsp__thread_inst_executed_per_inst_executed.ratio"][1])
    issue = float(m["smsp__issue_active.avg.pct_of_peak_sustained_active"][1])
    print(f"\nWarp execution efficiency {lanes:.2f}/32 = **{100 * lanes / 32:.1f} %**; issue slots "
          f"{issue:.1f} % busy; executed lane-instructions = issue x lanes = "
          f"**{issue / 100 * lanes / 32:.3f}** of one lane-instruction per lane, cycle and scheduler.\n")
    print("## Warp stall reasons (warps stalled per issue-active cycle)\n\n| reason | ratio |\n|---|---:|")
    st = [(re.match(STALLS, k).group(1), float(v[1])) for k, v in m.items() if re.match(STALLS, k)]
    for r, v in sorted(st, key=lambda x: -x[1]):
        if v >= 0.01:
            print(f"| {r} | {v:.3f} |")

    # ---- source view: rows with a line number are per-source-line aggregates
    src = ncu("-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass")
    cur_file, hdr = "?", None
    lines: dict[tuple[str, int], list] = {}
    for r in src:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].rsplit("/", 1)[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            ii, it, ip, ism = (hdr.index(c) for c in ("Instructions Executed", "Thread Instructions Executed",
                                                      "Predicated-On Thread Instructions Executed", "# Samples"))
            continue
        if hdr is None or len(r) != len(hdr) or not r[0].isdigit():
            continue
        lines[(cur_file, int(r[0]))] = [r[1], int(r[ii] or 0), int(r[it] or 0), int(r[ip] or 0), int(r[ism] or 0)]
    marks = sorted((ln, re.search(r"// =+ ([A-Z]): ", v[0]).group(1)) for (f, ln), v in lines.items()
                   if f == "play.cuh" and re.search(r"// =+ ([A-Z]): ", v[0]))
    if not marks:  # marker comment lines carry no instructions: read them from the working tree copy
        from pathlib import Path
        text = (Path(__file__).resolve().parents[1] / "farkle_ii_b200/csrc/play.cuh").read_text().splitlines()
        marks = [(i + 1, mm.group(1)) for i, l in enumerate(text) if (mm := re.search(r"// =+ ([A-Z]): ", l))]
        tail = next((i + 1 for i, l in enumerate(text) if "work counters: warp shuffle" in l), 10**9)
    else:
        tail = 10**9

    def phase(f: str, ln: int) -> str:
        if f != "play.cuh":
            return f
        if ln >= tail:
            return "Z totals"
        cur = "0 start_turn / setup (code above the loop)"
        for start, name in marks:
            if ln >= start:
                cur = name
        return cur

    agg: dict[str, list[int]] = defaultdict(lambda: [0, 0, 0])
    for (f, ln), v in lines.items():
        a = agg[phase(f, ln)]
        a[0] += v[1]; a[1] += v[2]; a[2] += v[4]
    ti = sum(a[0] for a in agg.values()); ts = max(sum(a[2] for a in agg.values()), 1)
    print(f"\n## Instruction breakdown by loop phase ({rolls:,.0f} lane-rolls in the launch)\n\n```")
    print(f"total warp-inst {ti:,}; per warp-iteration (32 lane-rolls): {ti * 32 / rolls:.1f}")
    for g, (i, t, s) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if i:
            print(f"{g:44s} warp-inst/iter {i * 32 / rolls:7.1f} ({100 * i / ti:5.1f}%)  thread-inst/roll "
                  f"{t / rolls:6.1f}  thr/inst {t / max(i, 1):5.1f}  samples {100 * s / ts:5.1f}%")
    print("```\n\n## Hottest source lines (warp instructions executed)\n\n```")
    for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"{f:12s}:{ln:4d} inst {100 * v[1] / ti:5.2f}%  thr/inst {v[2] / max(v[1], 1):5.1f}  pred-on "
              f"{v[3] / max(v[1], 1):5.1f}  samples {100 * v[4] / ts:5.2f}% | {v[0].strip()[:90]}")
    print("```")


if __name__ == "__main__":
    main()
