"""Binary layouts shared with ``include/farkle_b200.h`` (strategy entries, compact rows, tallies)."""

from __future__ import annotations

import numpy as np

TALLY_WIDTH = 26
LAG_WIDTH = 11  # pairs, win {sx, sy, sx2, sy2, sxy}, n_rounds {sx, sy, sx2, sy2, sxy}
MATCHUP_LAG_WIDTH = 6  # pairs, n_rounds {sx, sy, sx2, sy2, sxy}
ALLP_WIDTH = 45  # FB_ALLP_WIDTH: all-player statistics per (batch, strategy)
SEAT_TALLY_WIDTH = 4  # raw_wins, raw_exposures, raw_completed_exposures, raw_safety_limit_exposures
N_METRICS = 11
MAX_PLAYERS = 12
MAX_ROUNDS = 32767  # FB_MAX_ROUNDS: n_rounds is an int16 column (utils/schema_helpers.py:23-42)
TOTALS_WIDTH = 8 + MAX_PLAYERS

# tally columns (run_tournament.py:177-195, 109-121)
T_WINS, T_ATTEMPTED, T_COMPLETED, T_SAFETY, T_SUMS, T_SQ_SUMS = 0, 1, 2, 3, 4, 15
# totals columns
TOT_ATTEMPTED, TOT_COMPLETED, TOT_SAFETY, TOT_ROLLS, TOT_DICE, TOT_WORDS, TOT_TURNS, TOT_ERRORS = (
    range(8))
TOT_SEAT_WINS = 8

ROW_SAFETY_LIMIT = 0x01
ROW_ROLL_LIMIT = 0x02
ROW_I16_OVERFLOW = 0x04

STRATEGY_DTYPE = np.dtype(
    [("score_threshold", "<i4"), ("dice_threshold", "<i2"), ("flags", "<u2")]
)
assert STRATEGY_DTYPE.itemsize == 8

SF_SMART_FIVE = 0x01
SF_SMART_ONE = 0x02
SF_CONSIDER_SCORE = 0x04
SF_CONSIDER_DICE = 0x08
SF_REQUIRE_BOTH = 0x10
SF_AUTO_HOT_DICE = 0x20
SF_RUN_UP_SCORE = 0x40
SF_FAVOR_SCORE = 0x80

SEAT_DTYPE = np.dtype(
    [
        ("score", "<i4"),
        ("strategy", "<i4"),
        ("highest_turn", "<i4"),
        ("farkles", "<u2"),
        ("rolls", "<u2"),
        ("n_turns", "<u2"),
        ("hot_dice", "<u2"),
        ("smart_five_uses", "<u2"),
        ("n_smart_five_dice", "<u2"),
        ("smart_one_uses", "<u2"),
        ("n_smart_one_dice", "<u2"),
    ]
)
assert SEAT_DTYPE.itemsize == 28


def row_stride(k: int) -> int:
    """Bytes of one compact row (``fb_row_stride``)."""
    return (16 + 28 * k + 15) & ~15


def row_dtype(k: int) -> np.dtype:
    """Structured dtype of ``fb_row_header_t`` + k x ``fb_row_seat_t`` (padded)."""
    return np.dtype(
        {
            "names": ["game_seed", "game_ordinal", "n_rounds", "winner_seat", "flags", "seats"],
            "formats": ["<u8", "<u4", "<u2", "u1", "u1", (SEAT_DTYPE, (k,))],
            "offsets": [0, 8, 12, 14, 15, 16],
            "itemsize": row_stride(k),
        }
    )
