"""Strategy objects and grid enumeration (host side).

Mirror of the reference's ``farkle.simulation.strategies`` /
``farkle.simulation.simulation.generate_strategy_grid`` for the parts the hot
path needs: the ten-parameter ``ThresholdStrategy`` (strategies.py:165-290), the
grid order that defines strategy ids (``iter_strategy_combos`` strategies.py:346-396,
``_favor_options`` :335-343, inactive-threshold sentinels :519-535, stop-at
strategies :455-482) and the packing of a strategy list into the 8-byte
``fb_strategy_t`` table the kernels read.
"""

from __future__ import annotations

import re
from dataclasses import dataclass
from enum import Enum
from itertools import product
from typing import Any, Iterable, Iterator, Sequence

import numpy as np

from .layout import (
    SF_AUTO_HOT_DICE,
    SF_CONSIDER_DICE,
    SF_CONSIDER_SCORE,
    SF_FAVOR_SCORE,
    SF_REQUIRE_BOTH,
    SF_RUN_UP_SCORE,
    SF_SMART_FIVE,
    SF_SMART_ONE,
    STRATEGY_DTYPE,
)


class FavorDiceOrScore(Enum):
    """Tie-break preference of the smart-discard search."""

    SCORE = "score"
    DICE = "dice"

    def __str__(self) -> str:
        return self.value


STOP_AT_THRESHOLDS: tuple[int, ...] = (350, 400, 450, 500)

STRATEGY_TUPLE_FIELDS: tuple[str, ...] = (
    "score_threshold", "dice_threshold", "smart_five", "smart_one", "consider_score",
    "consider_dice", "require_both", "auto_hot_dice", "run_up_score", "favor_dice_or_score",
)

DEFAULT_STRATEGY_GRID: dict[str, tuple[object, ...]] = {
    "score_thresholds": tuple(range(200, 1400, 50)),
    "dice_thresholds": tuple(range(0, 5)),
    "smart_five_opts": (True, False),
    "smart_one_opts": (True, False),
    "consider_score_opts": (True, False),
    "consider_dice_opts": (True, False),
    "auto_hot_dice_opts": (False, True),
    "run_up_score_opts": (True, False),
}


def decide_continue(turn_score: int, dice_left: int, score_threshold: int, dice_threshold: int,
                    consider_score: bool, consider_dice: bool, require_both: bool) -> bool:
    """Keep rolling? (reference ``_decide_continue``, strategies.py:125-162)."""
    want_score = consider_score and turn_score < score_threshold
    want_dice = consider_dice and dice_left > dice_threshold
    if consider_score and consider_dice:
        return (want_score or want_dice) if require_both else (want_score and want_dice)
    if consider_score:
        return want_score
    return want_dice if consider_dice else False


@dataclass
class ThresholdStrategy:
    """Threshold-based keep/bank rule with the reference's field names and checks."""

    score_threshold: int = 300
    dice_threshold: int = 2
    smart_five: bool = False
    smart_one: bool = False
    consider_score: bool = True
    consider_dice: bool = True
    require_both: bool = False
    auto_hot_dice: bool = False
    run_up_score: bool = False
    favor_dice_or_score: FavorDiceOrScore = FavorDiceOrScore.SCORE
    strategy_id: int | None = None

    def __post_init__(self) -> None:
        if self.smart_one and not self.smart_five:
            raise ValueError("ThresholdStrategy: smart_one=True requires smart_five=True")
        if self.require_both and not (self.consider_score and self.consider_dice):
            raise ValueError(
                "ThresholdStrategy: require_both=True requires both "
                "consider_score=True and consider_dice=True")

    def decide(self, *, turn_score: int, dice_left: int, has_scored: bool, score_needed: int = 0,
               final_round: bool = False, score_to_beat: int = 0, running_total: int = 0) -> bool:
        """Host restatement of strategies.py:212-275 (the kernel evaluates the same rule)."""
        del score_needed
        if not has_scored and turn_score < 500:
            return True
        if final_round:
            if running_total <= score_to_beat:
                return True
            if not self.run_up_score:
                return False
        return decide_continue(turn_score, dice_left, self.score_threshold, self.dice_threshold,
                               self.consider_score, self.consider_dice, self.require_both)

    def __str__(self) -> str:
        favor = "FS" if self.favor_dice_or_score is FavorDiceOrScore.SCORE else "FD"
        return (
            f"Strat({self.score_threshold},{self.dice_threshold})"
            f"[{'S' if self.consider_score else '-'}{'D' if self.consider_dice else '-'}]"
            f"[{'F' if self.smart_five else '-'}{'O' if self.smart_one else '-'}{favor}]"
            f"[{'AND' if self.require_both else 'OR'}]"
            f"[{'H' if self.auto_hot_dice else '-'}{'R' if self.run_up_score else '-'}]"
        )


@dataclass
class StopAtStrategy(ThresholdStrategy):
    """Named ``stop_at_<n>[_heuristic]`` strategy (strategies.py:293-307)."""

    label: str = ""
    heuristic: bool = False

    def __post_init__(self) -> None:
        super().__post_init__()
        if not re.fullmatch(r"stop_at_\d+(?:_heuristic)?", self.label):
            raise ValueError(f"Invalid stop-at strategy label: {self.label!r}")

    def __str__(self) -> str:
        return self.label


def build_stop_at_strategy(threshold: int, *, heuristic: bool = False,
                           inactive_dice_threshold: int | None = None) -> StopAtStrategy:
    if threshold not in STOP_AT_THRESHOLDS:
        raise ValueError(f"Unregistered stop-at threshold: {threshold}")
    return StopAtStrategy(
        score_threshold=threshold,
        dice_threshold=-1 if inactive_dice_threshold is None else inactive_dice_threshold,
        smart_five=heuristic, smart_one=heuristic, consider_score=True, consider_dice=False,
        require_both=False, auto_hot_dice=heuristic, run_up_score=False,
        favor_dice_or_score=FavorDiceOrScore.SCORE,
        label=f"stop_at_{threshold}" + ("_heuristic" if heuristic else ""), heuristic=heuristic)


_STRAT_RE = re.compile(
    r"Strat\(\s*(\d+)\s*,\s*(\d+)\s*\)\[([S\-])([D\-])\]\[([F\-])([O\-])(FS|FD)\]\[(AND|OR)\]"
    r"\[([H\-])([R\-])\]")


def parse_strategy(text: str) -> ThresholdStrategy:
    """Inverse of ``ThresholdStrategy.__str__`` (strategies.py:846-887)."""
    m = _STRAT_RE.fullmatch(text)
    if not m:
        raise ValueError(f"Cannot parse strategy string: {text!r}")
    st, dt, cs, cd, sf, so, fav, rb, hd, rs = m.groups()
    return ThresholdStrategy(
        score_threshold=int(st), dice_threshold=int(dt), smart_five=sf == "F", smart_one=so == "O",
        consider_score=cs == "S", consider_dice=cd == "D", require_both=rb == "AND",
        auto_hot_dice=hd == "H", run_up_score=rs == "R",
        favor_dice_or_score=FavorDiceOrScore.SCORE if fav == "FS" else FavorDiceOrScore.DICE)


def strategy_tuple(strategy: ThresholdStrategy) -> tuple:
    return tuple(getattr(strategy, name) for name in STRATEGY_TUPLE_FIELDS)


def _favor_choices(smart_five: bool, consider_score: bool, consider_dice: bool):
    if consider_score and consider_dice:
        return ((FavorDiceOrScore.SCORE, FavorDiceOrScore.DICE) if smart_five
                else (FavorDiceOrScore.SCORE,))
    if consider_dice and not consider_score:
        return (FavorDiceOrScore.DICE,)
    return (FavorDiceOrScore.SCORE,)


def iter_strategy_combos(*, score_thresholds: Sequence[int], dice_thresholds: Sequence[int],
                         smart_five_opts: Sequence[bool], smart_one_opts: Sequence[bool],
                         consider_score_opts: Sequence[bool], consider_dice_opts: Sequence[bool],
                         auto_hot_dice_opts: Sequence[bool], run_up_score_opts: Sequence[bool],
                         inactive_score_threshold: int, inactive_dice_threshold: int,
                         allowed_smart_pairs: set[tuple[bool, bool]] | None = None
                         ) -> Iterator[tuple]:
    """Yield strategy tuples in the reference's id order.

    Nesting, outermost first: smart_five, smart_one, consider_score, consider_dice,
    score threshold, dice threshold, auto_hot_dice, run_up_score, require_both, favor.
    """
    for sf in smart_five_opts:
        for so in smart_one_opts:
            if so and not sf:
                continue
            if allowed_smart_pairs is not None and (sf, so) not in allowed_smart_pairs:
                continue
            for cs, cd in product(consider_score_opts, consider_dice_opts):
                scores = score_thresholds if cs else (inactive_score_threshold,)
                dices = dice_thresholds if cd else (inactive_dice_threshold,)
                both = (True, False) if (cs and cd) else (False,)
                for st, dt, hd, rs, rb, fav in product(scores, dices, auto_hot_dice_opts,
                                                       run_up_score_opts, both,
                                                       _favor_choices(sf, cs, cd)):
                    yield (int(st), int(dt), bool(sf), bool(so), bool(cs), bool(cd), bool(rb),
                           bool(hd), bool(rs), fav)


def _options(values, default, *, sort: bool) -> tuple:
    if values is None:
        return tuple(default)
    out = tuple(values)
    if sort and not isinstance(values, tuple):
        try:
            return tuple(sorted(out))
        except TypeError:
            return out
    return out


@dataclass(frozen=True)
class StrategyGridOptions:
    """Normalised grid inputs (strategies.py:504-616): lists are sorted, tuples kept."""

    score_thresholds: tuple[int, ...]
    dice_thresholds: tuple[int, ...]
    smart_five_opts: tuple[bool, ...]
    smart_one_opts: tuple[bool, ...]
    consider_score_opts: tuple[bool, ...]
    consider_dice_opts: tuple[bool, ...]
    auto_hot_dice_opts: tuple[bool, ...]
    run_up_score_opts: tuple[bool, ...]
    include_stop_at: bool = False
    include_stop_at_heuristic: bool = False

    @property
    def inactive_score_threshold(self) -> int:
        return min(self.score_thresholds) - 1

    @property
    def inactive_dice_threshold(self) -> int:
        return min(self.dice_thresholds) - 1

    @classmethod
    def from_inputs(cls, *, score_thresholds=None, dice_thresholds=None, smart_five_opts=None,
                    smart_one_opts=None, consider_score_opts=(True, False),
                    consider_dice_opts=(True, False), auto_hot_dice_opts=(False, True),
                    run_up_score_opts=(True, False), include_stop_at: bool = False,
                    include_stop_at_heuristic: bool = False) -> "StrategyGridOptions":
        d = DEFAULT_STRATEGY_GRID
        return cls(
            _options(score_thresholds, d["score_thresholds"], sort=True),
            _options(dice_thresholds, d["dice_thresholds"], sort=True),
            _options(smart_five_opts, d["smart_five_opts"], sort=True),
            _options(smart_one_opts, d["smart_one_opts"], sort=True),
            _options(consider_score_opts, d["consider_score_opts"], sort=True),
            _options(consider_dice_opts, d["consider_dice_opts"], sort=True),
            _options(auto_hot_dice_opts, d["auto_hot_dice_opts"], sort=True),
            _options(run_up_score_opts, d["run_up_score_opts"], sort=True),
            include_stop_at, include_stop_at_heuristic)

    def combos(self) -> list[tuple]:
        return list(iter_strategy_combos(
            score_thresholds=self.score_thresholds, dice_thresholds=self.dice_thresholds,
            smart_five_opts=self.smart_five_opts, smart_one_opts=self.smart_one_opts,
            consider_score_opts=self.consider_score_opts,
            consider_dice_opts=self.consider_dice_opts,
            auto_hot_dice_opts=self.auto_hot_dice_opts, run_up_score_opts=self.run_up_score_opts,
            inactive_score_threshold=self.inactive_score_threshold,
            inactive_dice_threshold=self.inactive_dice_threshold))


def generate_strategy_grid(*, score_thresholds=None, dice_thresholds=None, smart_five_opts=None,
                           smart_one_opts=None, consider_score_opts=(True, False),
                           consider_dice_opts=(True, False), auto_hot_dice_opts=(False, True),
                           run_up_score_opts=(True, False), include_stop_at: bool = False,
                           include_stop_at_heuristic: bool = False):
    """Build the strategy grid (reference ``generate_strategy_grid``, simulation.py:55-221).

    Returns ``(strategies, meta)``; ``meta`` is a pandas DataFrame with the reference's
    columns (parameter columns, ``strategy_id``, ``strategy_idx``).  A strategy's id is the
    position of its first occurrence in the enumeration (the reference's encoder).
    """
    import pandas as pd

    opts = StrategyGridOptions.from_inputs(
        score_thresholds=score_thresholds, dice_thresholds=dice_thresholds,
        smart_five_opts=smart_five_opts, smart_one_opts=smart_one_opts,
        consider_score_opts=consider_score_opts, consider_dice_opts=consider_dice_opts,
        auto_hot_dice_opts=auto_hot_dice_opts, run_up_score_opts=run_up_score_opts,
        include_stop_at=include_stop_at, include_stop_at_heuristic=include_stop_at_heuristic)
    if not opts.score_thresholds:
        raise ValueError("score_thresholds must contain at least one value")
    if not opts.dice_thresholds:
        raise ValueError("dice_thresholds must contain at least one value")
    combos = opts.combos()
    extra: list[StopAtStrategy] = []
    for flag, heuristic in ((opts.include_stop_at, False), (opts.include_stop_at_heuristic, True)):
        if flag:
            extra += [build_stop_at_strategy(t, heuristic=heuristic,
                                             inactive_dice_threshold=opts.inactive_dice_threshold)
                      for t in STOP_AT_THRESHOLDS]
    ids: dict[tuple, int] = {}
    for combo in [*combos, *(strategy_tuple(s) for s in extra)]:
        ids.setdefault(combo, len(ids))
    strategies: list[ThresholdStrategy] = [
        ThresholdStrategy(*combo, strategy_id=ids[combo]) for combo in combos]
    for s in extra:
        s.strategy_id = ids[strategy_tuple(s)]
        strategies.append(s)
    meta = pd.DataFrame([strategy_tuple(s) for s in strategies],
                        columns=list(STRATEGY_TUPLE_FIELDS))
    meta["strategy_id"] = [s.strategy_id for s in strategies]
    meta["strategy_idx"] = meta.index
    return strategies, meta


def prepare_strategy_ids(strategies: Sequence[ThresholdStrategy]) -> list[int]:
    """Resolve one unique id per seat position (``_prepare_public_helper_strategies``,
    simulation.py:361-409): keep given ids, fill gaps with the smallest unused integers."""
    given: list[int | None] = []
    used: set[int] = set()
    for pos, s in enumerate(strategies):
        sid = s.strategy_id
        if sid is None:
            given.append(None)
            continue
        if isinstance(sid, bool) or not isinstance(sid, (int, np.integer)) or int(sid) < 0:
            raise ValueError(f"strategies[{pos}].strategy_id must be a non-negative integer")
        if int(sid) in used:
            raise ValueError(f"Caller-provided strategy IDs must be unique; found {int(sid)}")
        used.add(int(sid))
        given.append(int(sid))
    out: list[int] = []
    nxt = 0
    for sid in given:
        if sid is None:
            while nxt in used:
                nxt += 1
            used.add(nxt)
            sid = nxt
            nxt += 1
        out.append(sid)
    return out


def pack_strategy(strategy: ThresholdStrategy) -> tuple[int, int, int]:
    """``(score_threshold, dice_threshold, flags)`` of one ``fb_strategy_t``."""
    flags = 0
    for attr, bit in (("smart_five", SF_SMART_FIVE), ("smart_one", SF_SMART_ONE),
                      ("consider_score", SF_CONSIDER_SCORE), ("consider_dice", SF_CONSIDER_DICE),
                      ("require_both", SF_REQUIRE_BOTH), ("auto_hot_dice", SF_AUTO_HOT_DICE),
                      ("run_up_score", SF_RUN_UP_SCORE)):
        if getattr(strategy, attr):
            flags |= bit
    favor = strategy.favor_dice_or_score
    if favor is True or favor is FavorDiceOrScore.SCORE:
        flags |= SF_FAVOR_SCORE
    st, dt = int(strategy.score_threshold), int(strategy.dice_threshold)
    if not (-2**31 <= st < 2**31 and -2**15 <= dt < 2**15):
        raise ValueError("strategy thresholds outside the int32 / int16 table range")
    return st, dt, flags


def pack_strategies(strategies: Iterable[ThresholdStrategy]) -> np.ndarray:
    """Pack strategies into the ``fb_strategy_t`` table (STRATEGY_DTYPE array)."""
    return np.array([pack_strategy(s) for s in strategies], dtype=STRATEGY_DTYPE)


def unpack_strategy(entry: Any, strategy_id: int | None = None) -> ThresholdStrategy:
    flags = int(entry["flags"])
    return ThresholdStrategy(
        score_threshold=int(entry["score_threshold"]), dice_threshold=int(entry["dice_threshold"]),
        smart_five=bool(flags & SF_SMART_FIVE), smart_one=bool(flags & SF_SMART_ONE),
        consider_score=bool(flags & SF_CONSIDER_SCORE), consider_dice=bool(flags & SF_CONSIDER_DICE),
        require_both=bool(flags & SF_REQUIRE_BOTH), auto_hot_dice=bool(flags & SF_AUTO_HOT_DICE),
        run_up_score=bool(flags & SF_RUN_UP_SCORE),
        favor_dice_or_score=(FavorDiceOrScore.SCORE if flags & SF_FAVOR_SCORE
                             else FavorDiceOrScore.DICE),
        strategy_id=strategy_id)
