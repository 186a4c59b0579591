#!/usr/bin/env python
"""Run the integer-issue probe variants: python scripts/issue_peak.py [VARIANT ...]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from farkle_ii_b200.device import get_engine  # noqa: E402

eng = get_engine(0)
names = ["chain-major 8", "op-major 8", "op-major 16 interleaved", "xor/add = LOP3 + IMAD.IADD", "mad only"]
nominal = eng.sm_count * 4 * 32 * 1.965e9
for v in ([int(a) for a in sys.argv[1:]] or range(5)):
    r = eng.measure_issue_peak_variant(v)
    print(f"variant {v} ({names[v]}): {r / 1e12:.2f} T lane-op/s = {r / nominal:.3f} of nominal "
          f"{nominal / 1e12:.1f} T ({eng.sm_count} SMs x 4 x 32 x 1.965 GHz)")
