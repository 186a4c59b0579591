"""Test double: an ``Engine`` look-alike whose compute calls go to the CPU oracle.

Only ``tests/`` may use this.  It lets the host-surface mirror (row expansion, counters,
H2H loop, rank sharding) be checked against the reference's golden outputs on a box without
a GPU; the ``-m gpu`` variants of the same tests run the real ``Engine``.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

import oracle
from farkle_ii_b200.layout import STRATEGY_DTYPE, row_dtype

P_H2H_PLAYER = 203


@dataclass
class _Result:
    tallies: torch.Tensor | None
    totals: torch.Tensor
    rows: torch.Tensor | None
    n_games: int
    k: int
    seat_tallies: torch.Tensor | None = None
    lag_stats: torch.Tensor | None = None
    lag_edges: torch.Tensor | None = None
    first_seen: torch.Tensor | None = None
    matchup_participants: torch.Tensor | None = None
    matchup_count: torch.Tensor | None = None
    matchup_stats: torch.Tensor | None = None

    def rows_numpy(self) -> np.ndarray:
        return self.rows.numpy().view(row_dtype(self.k)).reshape(-1)


def _table(strategies) -> np.ndarray:
    if isinstance(strategies, torch.Tensor):
        return strategies.cpu().numpy().view(STRATEGY_DTYPE).reshape(-1)
    return np.ascontiguousarray(strategies, dtype=STRATEGY_DTYPE)


class OracleEngine:
    device = torch.device("cpu")

    def to_device(self, a: np.ndarray, dtype=None) -> torch.Tensor:
        return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy())

    def play_tournament(self, root_seed, k, shuffle0, n_shuffles, strategies, *, strategy_ids=None,
                        n_tally_ids=None, target_score=10_000, max_rounds=200, overrides=(),
                        shuffles_per_slot=0, want_tallies=True, want_rows=False,
                        want_game_seeds=False, tallies=None, totals=None, lags=(),
                        want_first_seen=False, matchup_min_observations=0, strategy_lags=True):
        if want_first_seen and not lags:
            res = self.play_tournament(root_seed, k, shuffle0, n_shuffles, strategies,
                                       strategy_ids=strategy_ids, n_tally_ids=n_tally_ids,
                                       target_score=target_score, max_rounds=max_rounds, overrides=overrides,
                                       shuffles_per_slot=shuffles_per_slot, want_tallies=want_tallies,
                                       want_rows=True, want_game_seeds=want_game_seeds, tallies=tallies,
                                       totals=totals)
            res.first_seen = torch.from_numpy(self._first_seen(res.rows_numpy(), len(_table(strategies)), k,
                                                               strategy_ids))
            if not want_rows:
                res.rows = None
            return res
        if lags:
            res = self._with_lags(root_seed, k, shuffle0, n_shuffles, strategies, target_score, max_rounds,
                                  tuple(lags))
            if matchup_min_observations > 0:                  # the slow way: group the rows on the host
                from farkle_ii_b200.rng_diagnostics import MatchupLagGroups

                _, _, rows = oracle.play_tournament(root_seed, k, shuffle0, n_shuffles, _table(strategies),
                                                    strategy_ids=strategy_ids, target_score=target_score,
                                                    max_rounds=max_rounds, want_rows=True, n_threads=2)
                groups = MatchupLagGroups.from_rows(tuple(lags), rows, matchup_min_observations)
                res.matchup_participants = torch.from_numpy(groups.participants)
                res.matchup_count = torch.from_numpy(groups.counts.astype(np.int32))
                res.matchup_stats = torch.from_numpy(groups.stats)
            return res
        t, tot, rows = oracle.play_tournament(
            root_seed, k, shuffle0, n_shuffles, _table(strategies), strategy_ids=strategy_ids,
            n_tally_ids=n_tally_ids, target_score=target_score, max_rounds=max_rounds,
            overrides=list(overrides), shuffles_per_slot=shuffles_per_slot, want_rows=want_rows,
            want_game_seeds=want_game_seeds, n_threads=2)
        tt, to = torch.from_numpy(t), torch.from_numpy(tot)
        if tallies is not None:
            tallies += tt
            tt = tallies
        if totals is not None:
            totals += to
            to = totals
        r = None if rows is None else torch.from_numpy(rows.view(np.uint8).reshape(len(rows), -1))
        return _Result(tt if want_tallies else None, to, r, len(rows) if rows is not None else 0, k)

    @staticmethod
    def _first_seen(rows, n, k, strategy_ids):
        """First ordinal of: win, exposure, completed exposure, safety-limit exposure, per tally id."""
        ids = np.arange(n) if strategy_ids is None else np.asarray(strategy_ids)
        seen = np.full((int(ids.max()) + 1, 4), -1, dtype=np.int64)
        gps = n // k

        def note(sid, col, ordinal):
            if seen[sid, col] < 0:
                seen[sid, col] = ordinal
        for g, row in enumerate(rows):
            safety = bool(row["flags"] & 1)
            for s in range(k):
                sid = int(row["seats"]["strategy"][s])
                ordinal = (g // gps) * n + (g % gps) * k + s
                note(sid, 1, ordinal)
                note(sid, 3 if safety else 2, ordinal)
                if not safety and int(row["winner_seat"]) == s:
                    note(sid, 0, ordinal)
        return seen.astype(np.int32)

    def _with_lags(self, root_seed, k, shuffle0, n_shuffles, strategies, target_score, max_rounds, lags):
        """Lag sums the slow way: one accumulator walk per strategy over the oracle's rows."""
        table = _table(strategies)
        t, tot, rows = oracle.play_tournament(root_seed, k, shuffle0, n_shuffles, table,
                                              target_score=target_score, max_rounds=max_rounds,
                                              want_rows=True, n_threads=2)
        n, gps = len(table), len(table) // k
        seq = [[None] * n_shuffles for _ in range(n)]           # (win, rounds) per strategy and shuffle
        for g, row in enumerate(rows):
            safety = bool(row["flags"] & 1)
            for s in range(k):
                win = int((not safety) and int(row["winner_seat"]) == s)
                seq[int(row["seats"]["strategy"][s])][g // gps] = (win, int(row["n_rounds"]))
        stats = np.zeros((n, len(lags), 11), dtype=np.int64)
        max_lag = max(lags)
        edges = np.zeros((n, 2, max_lag), dtype=np.uint32)
        cnt = min(max_lag, n_shuffles)
        for i in range(n):
            for z, lag in enumerate(lags):
                for j in range(lag, n_shuffles):
                    (xw, xr), (yw, yr) = seq[i][j - lag], seq[i][j]
                    stats[i, z] += [1, xw, yw, xw * xw, yw * yw, xw * yw, xr, yr, xr * xr, yr * yr, xr * yr]
            for m in range(cnt):
                edges[i, 0, m] = seq[i][m][1] | seq[i][m][0] << 16
                last = seq[i][n_shuffles - cnt + m]
                edges[i, 1, m] = last[1] | last[0] << 16
        return _Result(torch.from_numpy(t), torch.from_numpy(tot), None, len(rows), k, None,
                       torch.from_numpy(stats), torch.from_numpy(edges.view(np.int32)))

    def play_games(self, coords, k, seat_strategies, **kw):
        return oracle.play_games(coords, k, seat_strategies, **kw)

    def play_h2h(self, root_seed, pair_id, order, seat1, seat2, attempt0, n_attempts, *,
                 target_score=10_000, max_rounds=200, want_rows=False):
        pair_id, order = np.asarray(pair_id), np.asarray(order)
        attempt0, n_attempts = np.asarray(attempt0), np.asarray(n_attempts, dtype=np.uint32)
        s1, s2 = _table(seat1), _table(seat2)
        coords, st = [], []
        for b in range(len(pair_id)):
            for a in range(int(attempt0[b]), int(attempt0[b]) + int(n_attempts[b])):
                coords.append([P_H2H_PLAYER, root_seed, 2, 0, int(pair_id[b]), int(order[b]), a])
                st.append([tuple(s1[b]), tuple(s2[b])])
        if not coords:
            return torch.zeros(0, dtype=torch.uint8), torch.from_numpy(n_attempts.astype(np.int64)), None, None
        rows, totals = oracle.play_games(np.array(coords, dtype=np.uint64), 2,
                                         np.array(st, dtype=STRATEGY_DTYPE), target_score=target_score,
                                         max_rounds=max_rounds)
        oc = np.where(rows["flags"] & 1, 0, rows["winner_seat"] + 1).astype(np.uint8)
        oc |= np.where(rows["flags"] & ~np.uint8(1), 0x80, 0).astype(np.uint8)
        return torch.from_numpy(oc), torch.from_numpy(n_attempts.astype(np.int64)), None, torch.from_numpy(totals)

    def h2h_resolve(self, d_n_attempts, outcome, required, progress):
        na = d_n_attempts.numpy()
        oc = outcome.numpy()
        prog = np.array(progress, dtype=np.int32).reshape(-1, 5)
        off = 0
        for b in range(len(na)):
            for o in oc[off:off + int(na[b])] & 0x7F:
                if prog[b, 1] >= required[b]:
                    break
                prog[b, 0] += 1
                if o == 0:
                    prog[b, 2] += 1
                else:
                    prog[b, 1] += 1
                    prog[b, 2 + int(o)] += 1
            off += int(na[b])
        return prog
