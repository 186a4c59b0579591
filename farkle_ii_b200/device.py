"""Device-side driver: torch owns memory and streams, the C ABI does the work.

PyTorch is plumbing here (allocation, streams, ``torch.distributed``); every
computation goes through ``libfarkle_b200.so``.  All methods raise
:class:`farkle_ii_b200._native.NativeError` if the library or a CUDA device is
missing — there is no CPU path.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _native
from .layout import (
    ALLP_WIDTH,
    LAG_WIDTH,
    MATCHUP_LAG_WIDTH,
    MAX_ROUNDS,
    SEAT_TALLY_WIDTH,
    STRATEGY_DTYPE,
    TALLY_WIDTH,
    TOTALS_WIDTH,
    row_dtype,
    row_stride,
)

_engines: dict[int, "Engine"] = {}


def get_engine(device: int | None = None) -> "Engine":
    """Return the (cached) engine bound to ``device`` (default: current CUDA device)."""
    if not torch.cuda.is_available():
        raise _native.NativeError(
            "no CUDA device is visible; farkle_ii_b200 has no CPU fallback")
    if device is None:
        device = torch.cuda.current_device()
    eng = _engines.get(device)
    if eng is None:      # one engine (and one library context) per device; several may coexist
        eng = _engines[device] = Engine(device)
    return eng


class _OnDevice:
    """The C library resolves its per-device context from the CUDA device that is current for the
    calling thread (include/farkle_b200.h, fb_init): every call of an engine goes through this
    proxy, which makes the engine's device current for the duration of the call."""

    def __init__(self, lib: C.CDLL, device: torch.device):
        self._lib, self._device = lib, device

    def __getattr__(self, name: str):
        fn = getattr(self._lib, name)
        device = self._device

        def call(*args):
            with torch.cuda.device(device):
                return fn(*args)

        call.__name__ = name
        setattr(self, name, call)
        return call


def _check_rounds(values) -> None:
    """Per-game ``max_rounds`` values obey the bound the C ABI enforces on its scalar argument
    (``FB_MAX_ROUNDS``: ``n_rounds`` is an int16 column; the reference fails at the Arrow
    conversion of such a row, simulation/simulation.py:566-573)."""
    v = np.asarray(values, dtype=np.int64)
    if v.size and int(v.max()) > MAX_ROUNDS:
        raise _native.NativeError(f"max_rounds={int(v.max())} above {MAX_ROUNDS}: n_rounds is an int16 column")


def _ptr(t: torch.Tensor | None) -> C.c_void_p | None:
    return None if t is None else C.c_void_p(t.data_ptr())


@dataclass
class TournamentResult:
    """Outputs of one ``fb_play_tournament`` launch (device tensors)."""

    tallies: torch.Tensor | None  # int64 [slots, ids, 26]
    totals: torch.Tensor          # int64 [20]
    rows: torch.Tensor | None     # uint8 [n_games, stride]
    n_games: int
    k: int
    seat_tallies: torch.Tensor | None = None  # int64 [slots, ids, k, 4]
    lag_stats: torch.Tensor | None = None     # int64 [n_strategies, n_lags, LAG_WIDTH]
    lag_edges: torch.Tensor | None = None     # int32 [n_strategies, 2, max_lag]: n_rounds | win << 16
    matchup_participants: torch.Tensor | None = None  # int32 [groups, k] sorted strategy ids
    matchup_count: torch.Tensor | None = None         # int32 [groups] games
    matchup_stats: torch.Tensor | None = None         # int64 [groups, n_lags, MATCHUP_LAG_WIDTH]
    first_seen: torch.Tensor | None = None    # int32 [ids, 4] first ordinal of win / seat / completed / safety
    all_player: torch.Tensor | None = None    # int64 [slots, ids, ALLP_WIDTH] (41..44: float64 bit patterns)

    def rows_numpy(self) -> np.ndarray:
        assert self.rows is not None
        return self.rows.cpu().numpy().view(row_dtype(self.k)).reshape(-1)


class Engine:
    """One CUDA device + the loaded C-ABI library."""

    def __init__(self, device: int):
        self.device_index = device
        self.device = torch.device("cuda", device)
        self.lib = _OnDevice(_native.lib(), self.device)
        with torch.cuda.device(self.device):
            _native.check(self.lib.fb_init(device))
        sm, khz, major, minor = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _native.check(self.lib.fb_device_info(C.byref(sm), C.byref(khz), C.byref(major),
                                              C.byref(minor)))
        self.sm_count, self.clock_khz = sm.value, khz.value
        self.compute_capability = (major.value, minor.value)
        self._ws: torch.Tensor | None = None
        self._cells_ws: torch.Tensor | None = None

    # ------------------------------------------------------------------ helpers
    def _stream(self) -> C.c_void_p:
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def to_device(self, a: np.ndarray, dtype=None) -> torch.Tensor:
        a = np.ascontiguousarray(a)
        if a.dtype.fields is not None or a.dtype in (np.uint16, np.uint32, np.uint64):
            t = torch.from_numpy(a.view(np.uint8).reshape(-1)).to(self.device)
        else:
            t = torch.from_numpy(a).to(self.device)
        return t

    def empty(self, n_bytes: int) -> torch.Tensor:
        return torch.empty(max(int(n_bytes), 1), dtype=torch.uint8, device=self.device)

    def workspace(self, n_bytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < n_bytes:
            self._ws = None
            self._ws = self.empty(n_bytes)
        return self._ws

    def workspace_bytes(self, k: int, n_games: int) -> int:
        return int(self.lib.fb_workspace_bytes(k, n_games))

    def last_play_kernel_ms(self) -> float:
        return float(self.lib.fb_last_play_kernel_ms())

    def play_kernel_ms_history(self, n: int) -> list[float]:
        """CUDA-event durations of the last ``n`` play_kernel launches of this thread, newest first."""
        buf = (C.c_float * max(n, 1))()
        got = self.lib.fb_play_kernel_ms_history(buf, n)
        if got < 0:
            _native.check(got)
        return [float(buf[i]) for i in range(got)]

    def measure_issue_peak(self, iters: int = 20000) -> float:
        """Measured 32-bit integer lane-instructions per second (roofline denominator)."""
        out = C.c_double()
        with torch.cuda.device(self.device):
            _native.check(self.lib.fb_measure_issue_peak(iters, C.byref(out)))
        return out.value

    def measure_issue_peak_variant(self, variant: int, iters: int = 20000) -> float:
        """One probe variant (see ``fb_measure_issue_peak_variant`` in include/farkle_b200.h)."""
        out = C.c_double()
        with torch.cuda.device(self.device):
            _native.check(self.lib.fb_measure_issue_peak_variant(variant, iters, C.byref(out)))
        return out.value

    def run_tournament_host(self, root_seed: int, k: int, shuffle0: int, n_shuffles: int,
                            strategies: np.ndarray, *, strategy_ids: np.ndarray | None = None,
                            n_tally_ids: int | None = None, target_score: int = 10_000,
                            max_rounds: int = 200, shuffles_per_slot: int = 0,
                            want_rows: bool = False, want_game_seeds: bool = False,
                            out_tallies: np.ndarray | None = None,
                            out_rows: np.ndarray | None = None):
        """HOST-buffer call (``fb_run_tournament_host``): H2D table, play, D2H tallies(/rows).

        Returns ``(tallies[slots, ids, 26], totals[20], rows | None)`` as host arrays.
        """
        st = np.ascontiguousarray(strategies, dtype=STRATEGY_DTYPE)
        n = len(st)
        ids = None if strategy_ids is None else np.ascontiguousarray(strategy_ids, dtype=np.int32)
        if n_tally_ids is None:
            n_tally_ids = n if ids is None else int(ids.max()) + 1
        n_slots = 1 if shuffles_per_slot <= 0 else -(-n_shuffles // shuffles_per_slot)
        tallies = out_tallies if out_tallies is not None else np.empty(
            (n_slots, n_tally_ids, TALLY_WIDTH), dtype=np.int64)
        totals = np.empty(TOTALS_WIDTH, dtype=np.int64)
        rows = None
        if want_rows:
            n_games = n_shuffles * (n // k)
            rows = out_rows if out_rows is not None else np.empty(n_games, dtype=row_dtype(k))
        vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)  # noqa: E731
        with torch.cuda.device(self.device):
            _native.check(self.lib.fb_run_tournament_host(
                root_seed, k, shuffle0, n_shuffles, vp(st), vp(ids), n, n_tally_ids, target_score,
                max_rounds, shuffles_per_slot, vp(tallies), vp(totals), vp(rows),
                int(want_game_seeds)))
        return tallies, totals, rows

    def kernel_launch_count(self) -> int:
        return int(self.lib.fb_kernel_launch_count())

    def timeline(self, enable: bool) -> None:
        """Start (and clear) or stop the per-kernel timeline log (``fb_timeline``)."""
        _native.check(self.lib.fb_timeline(int(enable)))

    def timeline_marks(self) -> list[tuple[int, str, float]]:
        """``(lane, mark, ms since the first mark)`` per recorded kernel; lane 1 is the preparation
        stream of ``play_cells``.  Waits for the recorded work."""
        buf = C.create_string_buffer(1 << 18)
        rc = self.lib.fb_timeline_dump(buf, len(buf))
        if rc < 0:
            _native.check(rc)
        out = []
        for line in buf.value.decode().splitlines():
            lane, name, ms = line.split()
            out.append((int(lane), name, float(ms)))
        return out

    # ----------------------------------------------------------- building blocks
    def seedseq_generate(self, entropy: np.ndarray, n_words: int) -> np.ndarray:
        e = np.ascontiguousarray(entropy, dtype=np.uint32)
        if e.ndim == 1:
            e = e[None, :]
        d_e = self.to_device(e)
        out = self.empty(e.shape[0] * n_words * 4)
        _native.check(self.lib.fb_seedseq_generate(_ptr(d_e), e.shape[1], e.shape[0], n_words,
                                                   _ptr(out), self._stream()))
        return out.cpu().numpy().view(np.uint32).reshape(e.shape[0], n_words)

    def coordinate_seeds(self, purpose: int, *, root_seed: int, k: int = 0, shuffle_index: int = 0,
                         pair_id: int = 0, order: int = 0, game_index: int = 0, vary: str,
                         base: int = 0, n: int, as_u32: bool = False) -> np.ndarray:
        vary_i = {"shuffle_index": 0, "game_index": 1, "pair_id": 2}[vary]
        out = self.empty(n * 8)
        _native.check(self.lib.fb_coordinate_seeds(purpose, root_seed, k, shuffle_index, pair_id,
                                                   order, game_index, vary_i, base, n,
                                                   int(as_u32), _ptr(out), self._stream()))
        return out.cpu().numpy().view(np.uint64)

    def seed_streams(self, coords: np.ndarray) -> np.ndarray:
        c = np.ascontiguousarray(coords, dtype=np.uint64).reshape(-1, 9)
        d_c = self.to_device(c)
        out = self.empty(len(c) * 32)
        _native.check(self.lib.fb_seed_streams(_ptr(d_c), len(c), _ptr(out), self._stream()))
        return out.cpu().numpy().view(np.uint64).reshape(-1, 4)

    def roll_dice(self, state_inc: np.ndarray, n_dice, half_buffer: np.ndarray | None = None
                  ) -> np.ndarray:
        si = np.ascontiguousarray(state_inc, dtype=np.uint64).reshape(-1, 4)
        nd = np.ascontiguousarray(n_dice, dtype=np.int32)
        d_si, d_nd = self.to_device(si), self.to_device(nd)
        d_hb = None
        if half_buffer is not None:
            d_hb = self.to_device(np.ascontiguousarray(half_buffer, dtype=np.uint32).reshape(-1, 2))
        out = self.empty(len(si) * len(nd) * 6)
        _native.check(self.lib.fb_roll_dice(_ptr(d_si), _ptr(d_hb), len(si), _ptr(d_nd), len(nd),
                                            _ptr(out), self._stream()))
        return out.cpu().numpy().reshape(len(si), len(nd), 6)

    def default_score(self, faces: np.ndarray, turn_score_pre, strategies: np.ndarray) -> np.ndarray:
        f = np.ascontiguousarray(faces, dtype=np.uint8).reshape(-1, 6)
        ts = np.ascontiguousarray(turn_score_pre, dtype=np.int32)
        st = np.ascontiguousarray(strategies, dtype=STRATEGY_DTYPE)
        assert len(f) == len(ts) == len(st)
        d_f, d_ts, d_st = self.to_device(f), self.to_device(ts), self.to_device(st)
        out = self.empty(len(f) * 20)
        _native.check(self.lib.fb_default_score(_ptr(d_f), _ptr(d_ts), _ptr(d_st), len(f),
                                                _ptr(out), self._stream()))
        return out.cpu().numpy().view(np.int32).reshape(-1, 5)

    def permute_shuffles(self, root_seed: int, k: int, shuffle0: int, n_shuffles: int,
                         n_strategies: int) -> np.ndarray:
        out = self.empty(n_shuffles * n_strategies * 4)
        _native.check(self.lib.fb_permute_shuffles(root_seed, k, shuffle0, n_shuffles,
                                                   n_strategies, _ptr(out), self._stream()))
        return out.cpu().numpy().view(np.int32).reshape(n_shuffles, n_strategies)

    # ------------------------------------------------------------------ hot path
    def play_tournament(self, root_seed: int, k: int, shuffle0: int, n_shuffles: int,
                        strategies: torch.Tensor | np.ndarray, *, strategy_ids=None,
                        n_tally_ids: int | None = None, target_score: int = 10_000,
                        max_rounds: int = 200, overrides=(), shuffles_per_slot: int = 0,
                        want_tallies: bool = True, want_rows: bool = False,
                        want_game_seeds: bool = False, tallies: torch.Tensor | None = None,
                        totals: torch.Tensor | None = None, want_seat_tallies: bool = False,
                        seat_tallies: torch.Tensor | None = None,
                        lags: tuple[int, ...] = (), matchup_min_observations: int = 0,
                        strategy_lags: bool = True, want_first_seen: bool = False,
                        want_all_player: bool = False) -> TournamentResult:
        """Enqueue shuffles ``shuffle0 .. shuffle0+n_shuffles-1`` of cell (root_seed, k).

        ``lags`` asks for the RNG lag statistics of the strategy groups (``lag_stats``
        int64 [n_strategies, n_lags, LAG_WIDTH] by table position, ``lag_edges`` uint32
        [n_strategies, 2, max(lags)]); ``matchup_min_observations`` > 0 adds the matchup
        groups with at least that many games (``matchup_participants`` / ``matchup_count`` /
        ``matchup_stats``); see ``farkle_ii_b200.rng_diagnostics``.  ``want_first_seen`` adds
        ``first_seen`` int32 [n_tally_ids, 4]: the first exposure ordinal at which an id won / was
        seated / was seated in a completed / in a safety-limit game (-1 = never), i.e. the key
        insertion order of the reference's counters.

        ``strategies`` is a device uint8 tensor holding ``fb_strategy_t`` entries, or a
        host STRATEGY_DTYPE array (copied).  ``tallies`` / ``totals`` may be passed to
        accumulate across launches.
        """
        if isinstance(strategies, np.ndarray):
            strategies = self.to_device(np.ascontiguousarray(strategies, dtype=STRATEGY_DTYPE))
        n_strategies = strategies.numel() // 8
        d_ids = None
        if strategy_ids is not None:
            d_ids = strategy_ids if isinstance(strategy_ids, torch.Tensor) else self.to_device(
                np.ascontiguousarray(strategy_ids, dtype=np.int32))
        if n_tally_ids is None:
            if strategy_ids is None:
                n_tally_ids = n_strategies
            else:
                n_tally_ids = int(np.asarray(strategy_ids).max()) + 1
        if want_tallies and n_tally_ids < 1:
            raise _native.NativeError(f"n_tally_ids={n_tally_ids} must be >= 1")
        gps = n_strategies // k if k > 0 else 0
        n_games = n_shuffles * gps
        n_slots = 1 if shuffles_per_slot <= 0 else -(-n_shuffles // shuffles_per_slot)
        if want_tallies and tallies is None:
            tallies = torch.zeros((n_slots, n_tally_ids, TALLY_WIDTH), dtype=torch.int64,
                                  device=self.device)
        if totals is None:
            totals = torch.zeros(TOTALS_WIDTH, dtype=torch.int64, device=self.device)
        if want_seat_tallies and seat_tallies is None:
            seat_tallies = torch.zeros((n_slots, n_tally_ids, max(k, 1), SEAT_TALLY_WIDTH),
                                       dtype=torch.int64, device=self.device)
        rows = None
        if want_rows:
            rows = torch.empty((max(n_games, 1), row_stride(k)), dtype=torch.uint8,
                               device=self.device)[:n_games]
        ov = list(overrides)
        _check_rounds([o[2] for o in ov])
        d_os = d_og = d_om = None
        if ov:
            d_os = self.to_device(np.array([o[0] for o in ov], dtype=np.uint64))
            d_og = self.to_device(np.array([o[1] for o in ov], dtype=np.uint32))
            d_om = self.to_device(np.array([o[2] for o in ov], dtype=np.int32))
        ws_bytes = self.workspace_bytes(max(k, 1), n_games) + 2 * (n_shuffles * n_strategies * 4 + 256)
        ws = self.workspace(ws_bytes)
        lag_stats = lag_edges = m_part = m_count = m_stats = first_seen = None
        request = None
        all_player = None
        if lags or want_first_seen or want_all_player:
            request = _native.LagRequest()
        if want_all_player:
            # unconditional all-player statistics per (deterministic batch, strategy):
            # analysis/all_player_metrics.py; see ``run_tournament.all_player_table``
            if shuffles_per_slot <= 0:
                raise _native.NativeError("want_all_player needs shuffles_per_slot > 0 (slot = deterministic batch)")
            all_player = torch.zeros((n_slots, n_tally_ids, ALLP_WIDTH), dtype=torch.int64, device=self.device)
            request.all_player_dev = all_player.data_ptr()
        if want_first_seen:
            first_seen = torch.empty((n_tally_ids, 4), dtype=torch.int32, device=self.device)
            request.first_seen_dev = first_seen.data_ptr()
        if lags:
            lags_c = (C.c_int32 * len(lags))(*[int(v) for v in lags])
            n_found = C.c_int64(0)
            request.lags = C.cast(lags_c, C.c_void_p)
            request.n_lags = len(lags)
            if strategy_lags:
                lag_stats = torch.zeros((n_strategies, len(lags), LAG_WIDTH), dtype=torch.int64,
                                        device=self.device)
                lag_edges = torch.zeros((n_strategies, 2, max(int(v) for v in lags)), dtype=torch.int32,
                                        device=self.device)
                request.strategy_stats_dev, request.strategy_edges_dev = lag_stats.data_ptr(), lag_edges.data_ptr()
            if matchup_min_observations > 0:
                capacity = max(n_games // matchup_min_observations, 1)
                m_part = torch.empty((capacity, max(k, 1)), dtype=torch.int32, device=self.device)
                m_count = torch.empty(capacity, dtype=torch.int32, device=self.device)
                m_stats = torch.empty((capacity, len(lags), MATCHUP_LAG_WIDTH), dtype=torch.int64,
                                      device=self.device)
                scratch = self.empty(int(self.lib.fb_matchup_scratch_bytes(n_games)) + 256)
                request.matchup_min_observations = int(matchup_min_observations)
                request.matchup_capacity = capacity
                request.matchup_participants_dev = m_part.data_ptr()
                request.matchup_count_dev, request.matchup_stats_dev = m_count.data_ptr(), m_stats.data_ptr()
                request.scratch_dev, request.scratch_bytes = scratch.data_ptr(), scratch.numel()
                request.n_matchups_host = C.cast(C.pointer(n_found), C.c_void_p)
        _native.check(self.lib.fb_play_tournament_lags(
            root_seed, k, shuffle0, n_shuffles, _ptr(strategies), _ptr(d_ids), n_strategies,
            n_tally_ids, target_score, max_rounds, _ptr(d_os), _ptr(d_og), _ptr(d_om), len(ov),
            shuffles_per_slot, _ptr(tallies if want_tallies else None), _ptr(totals), _ptr(rows),
            int(want_game_seeds), _ptr(seat_tallies), C.byref(request) if request is not None else None,
            _ptr(ws), ws.numel(), self._stream()))
        if m_part is not None:
            found = int(n_found.value)
            m_part, m_count, m_stats = m_part[:found], m_count[:found], m_stats[:found]
        return TournamentResult(tallies if want_tallies else None, totals, rows, n_games, k, seat_tallies,
                                lag_stats, lag_edges, m_part, m_count, m_stats, first_seen, all_player)

    def play_cells(self, cells, strategies: torch.Tensor | np.ndarray, *, ahead=None,
                   strategy_ids=None, n_tally_ids: int | None = None, target_score: int = 10_000,
                   max_rounds: int = 200, shuffles_per_slot: int = 0) -> None:
        """Enqueue several tournament cells, pipelined (``fb_play_tournament_cells``).

        ``cells``: sequence of ``(root_seed, k, shuffle0, n_shuffles, tallies, totals)`` with the
        int64 device tensors each cell accumulates into (``tallies`` [slots, ids, 26], ``totals``
        [20]; either may be ``None``).  ``ahead``: one more ``(root_seed, k, shuffle0, n_shuffles)``
        that is only prepared (permutations, seat seeding), so that the next call starting with
        exactly that cell plays at once.  Same results as one ``play_tournament`` per cell.
        """
        if isinstance(strategies, np.ndarray):
            strategies = self.to_device(np.ascontiguousarray(strategies, dtype=STRATEGY_DTYPE))
        n_strategies = strategies.numel() // 8
        d_ids = None
        if strategy_ids is not None:
            d_ids = strategy_ids if isinstance(strategy_ids, torch.Tensor) else self.to_device(
                np.ascontiguousarray(strategy_ids, dtype=np.int32))
        if n_tally_ids is None:
            n_tally_ids = n_strategies if strategy_ids is None else int(np.asarray(strategy_ids).max()) + 1
        every = list(cells) + ([tuple(ahead) + (None, None)] if ahead is not None else [])
        arr = (_native.Cell * max(len(every), 1))()
        keep = []
        for i, (root, k, s0, n, tallies, totals) in enumerate(every):
            arr[i].root_seed, arr[i].shuffle0, arr[i].k, arr[i].n_shuffles = int(root), int(s0), int(k), int(n)
            for name, t, shape_tail in (("tallies_dev", tallies, (n_tally_ids, TALLY_WIDTH)),
                                        ("totals_dev", totals, (TOTALS_WIDTH,))):
                if t is not None:
                    if t.dtype != torch.int64 or not t.is_contiguous() or tuple(t.shape[-len(shape_tail):]) != shape_tail:
                        raise _native.NativeError(f"cell {i}: {name} must be a contiguous int64 tensor [..., "
                                                  f"{', '.join(map(str, shape_tail))}]")
                    keep.append(t)
                setattr(arr[i], name, None if t is None else t.data_ptr())
        need = int(self.lib.fb_cells_workspace_bytes(arr, len(every), n_strategies))
        if need == 0:
            raise _native.NativeError("bad cell list (k must divide the strategy count)")
        if self._cells_ws is None or self._cells_ws.numel() < need:
            torch.cuda.synchronize(self.device)   # the library's preparation stream may still write the old one
            self._cells_ws = None
            self._cells_ws = self.empty(need)
        with torch.cuda.device(self.device):
            _native.check(self.lib.fb_play_tournament_cells(
                arr, len(every) - (1 if ahead is not None else 0), 1 if ahead is not None else 0,
                _ptr(strategies), _ptr(d_ids), n_strategies, n_tally_ids, target_score, max_rounds,
                shuffles_per_slot, _ptr(self._cells_ws), self._cells_ws.numel(), self._stream()))

    def play_games(self, coords: np.ndarray, k: int, seat_strategies: np.ndarray, *,
                   seat_strategy_ids=None, target_score: int = 10_000, max_rounds: int = 200,
                   target_scores=None, max_rounds_v=None) -> tuple[np.ndarray, np.ndarray]:
        """Games at explicit coordinates -> ``(rows, totals)`` as host arrays."""
        cc = np.ascontiguousarray(coords, dtype=np.uint64).reshape(-1, 7)
        n = len(cc)
        st = np.ascontiguousarray(seat_strategies, dtype=STRATEGY_DTYPE).reshape(n, k)
        d_cc, d_st = self.to_device(cc), self.to_device(st)
        d_ids = None if seat_strategy_ids is None else self.to_device(
            np.ascontiguousarray(seat_strategy_ids, dtype=np.int32).reshape(n, k))
        d_ts = None if target_scores is None else self.to_device(
            np.ascontiguousarray(target_scores, dtype=np.int32))
        if max_rounds_v is not None:
            _check_rounds(max_rounds_v)
        d_mr = None if max_rounds_v is None else self.to_device(
            np.ascontiguousarray(max_rounds_v, dtype=np.int32))
        rows = torch.empty((max(n, 1), row_stride(k)), dtype=torch.uint8, device=self.device)[:n]
        totals = torch.zeros(TOTALS_WIDTH, dtype=torch.int64, device=self.device)
        ws = self.workspace(self.workspace_bytes(k, n))
        _native.check(self.lib.fb_play_games(_ptr(d_cc), n, k, _ptr(d_st), _ptr(d_ids), _ptr(d_ts),
                                             target_score, _ptr(d_mr), max_rounds, _ptr(rows),
                                             _ptr(totals), _ptr(ws), ws.numel(), self._stream()))
        return (rows.cpu().numpy().view(row_dtype(k)).reshape(-1), totals.cpu().numpy())

    def play_h2h(self, root_seed: int, pair_id, order, seat1, seat2, attempt0, n_attempts, *,
                 target_score: int = 10_000, max_rounds: int = 200, want_rows: bool = False):
        """Play attempts ``attempt0[b] .. +n_attempts[b]`` of every block.

        Returns ``(outcome uint8 device tensor, d_n_attempts, rows | None, totals)``.
        """
        pair = np.ascontiguousarray(pair_id, dtype=np.uint64)
        nb = len(pair)
        na = np.ascontiguousarray(n_attempts, dtype=np.uint32)
        total = int(na.sum())
        d_pair = self.to_device(pair)
        d_order = self.to_device(np.ascontiguousarray(order, dtype=np.uint8))
        d_s1 = self.to_device(np.ascontiguousarray(seat1, dtype=STRATEGY_DTYPE))
        d_s2 = self.to_device(np.ascontiguousarray(seat2, dtype=STRATEGY_DTYPE))
        d_a0 = self.to_device(np.ascontiguousarray(attempt0, dtype=np.uint32))
        d_na = self.to_device(na)
        outcome = torch.empty(max(total, 1), dtype=torch.uint8, device=self.device)[:total]
        rows = None
        if want_rows:
            rows = torch.empty((max(total, 1), row_stride(2)), dtype=torch.uint8,
                               device=self.device)[:total]
        totals = torch.zeros(TOTALS_WIDTH, dtype=torch.int64, device=self.device)
        ws = self.workspace(self.workspace_bytes(2, total) + (nb + 1) * 8 + nb * 16 + 1024 + 512)
        _native.check(self.lib.fb_play_h2h(root_seed, nb, _ptr(d_pair), _ptr(d_order), _ptr(d_s1),
                                           _ptr(d_s2), _ptr(d_a0), _ptr(d_na), total, target_score,
                                           max_rounds, _ptr(outcome), _ptr(rows), _ptr(totals),
                                           _ptr(ws), ws.numel(), self._stream()))
        return outcome, d_na, rows, totals

    def h2h_resolve(self, d_n_attempts: torch.Tensor, outcome: torch.Tensor, required,
                    progress: np.ndarray) -> np.ndarray:
        """Apply the reference's early-stop rule; returns the updated ``progress[n, 5]``."""
        req = np.ascontiguousarray(required, dtype=np.int32)
        nb = len(req)
        d_req = self.to_device(req)
        d_prog = self.to_device(np.ascontiguousarray(progress, dtype=np.int32).reshape(nb, 5))
        _native.check(self.lib.fb_h2h_resolve(nb, _ptr(d_n_attempts), _ptr(outcome), _ptr(d_req),
                                              _ptr(d_prog), self._stream()))
        return d_prog.cpu().numpy().reshape(nb, 5)
