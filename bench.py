#!/usr/bin/env python
"""bench.py — simulated games/s of the tournament hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1], default_config as BASELINE.json describes it): the
full 5,160-strategy grid, k in [2, 4], roots [42, 43].  ONE STEP = one root's two
(root, k) cells at the planned 4,300 shuffles each (11,094,000 + 5,547,000 games):
permutations -> per-seat PCG64DXSM seeding -> whole games -> per-strategy tallies.  Step i
uses root 42 + (i % 2).  With N ranks (weak scaling) rank r plays shuffles
[r*4300, (r+1)*4300) of every cell — disjoint coordinate sub-streams — and the int64
tally tensors are merged with one NCCL all-reduce per cell, inside the timed region.

`value`  = games of all ranks / max-over-ranks CUDA-event time, strategy table resident
           in HBM, tallies left in HBM.
`e2e`    = the same cells through the HOST-buffer C-ABI call (fb_run_tournament_host via
           Engine.run_tournament_host): the strategy table is copied H2D and the
           tallies/totals D2H inside the timed region, every step.
The path is integer-issue bound: `roofline` reports algorithmic 32-bit
lane instructions per second of play_kernel against the measured issue peak
(fb_measure_issue_peak) and carries the HBM view as well.
`--impl reference` times the CPU restatement of the reference (oracle/, all host
threads) on a bounded sample of the same cells.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_STRATEGIES = 5160
SHUFFLES = 4300                      # workload_planner: delta 0.03 -> 4265 -> 43 x 100
SHUFFLES_PER_BATCH = 43
CELLS_K = (2, 4)
ROOTS = (42, 43)
MEGA_K, MEGA_ROOT = (2, 3, 4, 5, 6, 8, 10, 12), 102   # configs/farkle_mega_config.yaml:10-14
# Lane-instruction model of SURVEY.md §8d (play kernel): W per PCG64-DXSM word, D per die, R per roll.
# SURVEY's estimates were 28 / 8 / 90; the constants used are SASS-exact, from the ncu captures of the
# two play_kernel variants (scripts/sass_model.py -> profiles/sass_model.json; these defaults are the
# round-2 values).  With the face queue a die costs nothing beyond its share of a word and of the roll
# (D = 0): W words + R rolls, with the words the reference's generators would have produced, is the
# algorithmic work; E rolls is what the kernel executes (the top-up computes ~10 % more words than
# are consumed, and everything runs at the lane occupancy the kernel achieves).
SASS_MODEL_DEFAULT = {"k2": {"W": 32.17, "D": 0.0, "R": 139.71, "E": 210.92},
                      "generic": {"W": 32.4, "D": 0.0, "R": 148.86, "E": 222.65}}
SURVEY_MODEL = {"W": 28.0, "D": 8.0, "R": 90.0}


def sass_model() -> dict:
    try:
        got = json.loads((ROOT / "profiles" / "sass_model.json").read_text())
        return {kind: {c: float(got[kind][c]) for c in "WDRE"} | {"source": got[kind].get("source")}
                for kind in ("k2", "generic")}
    except Exception:
        return SASS_MODEL_DEFAULT
METRIC = "simulated games/sec at 1/2/4/8 B200 (bit-exact tallies) vs reference CPU n_jobs"


def full_grid_table() -> np.ndarray:
    from farkle_ii_b200.strategies import generate_strategy_grid, pack_strategies

    strategies, _ = generate_strategy_grid()
    assert len(strategies) == N_STRATEGIES
    return pack_strategies(strategies)


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.

    In-process NVML (nvidia_ml_py) every 20 ms: a handful of driver queries, no process start.
    Spawning `nvidia-smi` ten times a second, as the first version did, measurably slowed the
    kernels it was watching (k=4 cell 16.4 -> 16.8 ms); it remains the fallback when NVML cannot
    be loaded.
    """

    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples: list[tuple[int, int]] = []       # (sm MHz, reason bits in NAMES order)
        self.sm_max: int | None = None
        self.source = "nvml"
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._nvml = self._handle = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml = pynvml
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self._handle, pynvml.NVML_CLOCK_SM))
            self._masks = (pynvml.nvmlClocksEventReasonHwSlowdown, pynvml.nvmlClocksEventReasonHwThermalSlowdown,
                           pynvml.nvmlClocksEventReasonSwThermalSlowdown, pynvml.nvmlClocksEventReasonSwPowerCap)
        except Exception:
            self._nvml = None
            self.source = "nvidia-smi"

    def _sample_nvml(self) -> None:
        n = self._nvml
        sm = int(n.nvmlDeviceGetClockInfo(self._handle, n.NVML_CLOCK_SM))
        bits = int(n.nvmlDeviceGetCurrentClocksEventReasons(self._handle))
        self.samples.append((sm, sum(1 << i for i, m in enumerate(self._masks) if bits & m)))

    def _sample_smi(self) -> None:
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}",
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
        if out.returncode == 0 and out.stdout.strip():
            f = [x.strip() for x in out.stdout.strip().split(",")]
            if f[0].isdigit():
                self.sm_max = int(f[1]) if f[1].isdigit() else self.sm_max
                self.samples.append((int(f[0]), sum(1 << i for i in range(4)
                                                    if f[2 + i].lower().startswith("active"))))

    def _run(self) -> None:
        while not self._stop.is_set():
            try:
                self._sample_nvml() if self._nvml else self._sample_smi()
            except Exception:
                pass
            self._stop.wait(0.02 if self._nvml else 0.1)

    def __enter__(self):
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=6)

    def summary(self) -> dict:
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["unsampled"]}
        sm = sorted(s[0] for s in self.samples)
        seen = 0
        for _, bits in self.samples:
            seen |= bits
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.sm_max, "samples": len(self.samples),
                "reasons": [n for i, n in enumerate(self.NAMES) if seen >> i & 1], "source": self.source}


def lane_ops(totals: np.ndarray, c: dict) -> tuple[float, float]:
    """(algorithmic, executed) lane instructions of one play_kernel launch from its work counters
    (totals: 3 rolls, 4 dice, 5 rng words) and the model constants `c`."""
    rolls, dice, words = float(totals[3]), float(totals[4]), float(totals[5])
    executed = c["E"] * rolls if "E" in c else (3 * c["W"] + 6 * c["D"] + c["R"]) * rolls
    return c["W"] * words + c["D"] * dice + c["R"] * rolls, executed


# ----------------------------------------------------------------------------- CPU arm
def cpu_sample(table: np.ndarray, seconds_target: float, threads: int) -> dict:
    """Time the oracle (kind "port") on a bounded sample of the step's cells."""
    import oracle

    oracle.build()
    # calibrate on a small slice, then size the sample
    t0 = time.perf_counter()
    oracle.play_tournament(ROOTS[0], 2, 0, threads, table, n_threads=threads)
    per_shuffle = (time.perf_counter() - t0) / threads
    n_sh = int(max(threads, min(SHUFFLES, seconds_target / max(per_shuffle, 1e-6) / 1.5)))
    n_sh2 = max(threads, (n_sh * 2) // 3)
    n_sh4 = max(threads, n_sh - n_sh2)
    t0 = time.perf_counter()
    _, tot2, _ = oracle.play_tournament(ROOTS[0], 2, 0, n_sh2, table, n_threads=threads)
    _, tot4, _ = oracle.play_tournament(ROOTS[0], 4, 0, n_sh4, table, n_threads=threads)
    dt = time.perf_counter() - t0
    games = int(tot2[0] + tot4[0])
    return {"value": games / dt, "unit": "games/s", "cores": threads, "kind": "port",
            "sample": (f"full grid root {ROOTS[0]}: k=2 shuffles 0..{n_sh2 - 1} + k=4 shuffles "
                       f"0..{n_sh4 - 1} ({games} games, {dt:.1f} s), C restatement of the reference "
                       f"(oracle/farkle_oracle.c), {threads} pthreads over shuffles"),
            "seconds": dt, "games": games}


def python_reference_sample(k: int, root: int, shuffles: int, n_jobs: int, warm: bool = False) -> dict | None:
    """The UNMODIFIED reference's own pool path (`run_tournament.run_tournament`, n_jobs workers)
    timed on this box's host cores by scripts/time_reference.py in a process of its own, from the
    staged checkout baseline/_ref/ (scripts/stage_reference.sh).  None when it is not staged."""
    cmd = [sys.executable, str(ROOT / "scripts" / "time_reference.py"), "--k", str(k), "--root", str(root),
           "--shuffles", str(shuffles), "--n-jobs", str(n_jobs)] + (["--warm"] if warm else [])
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
        res = json.loads(line)
    except Exception as exc:  # the baseline is reported, never required
        return {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
    return None if "unavailable" in res else res


def python_reference_leg(threads: int) -> dict:
    """`cpu_baseline_reference`: J = 1 (one shuffle) and J = all host cores (one shuffle per core) of
    the k=2 full-grid cell, root 42 (SURVEY.md section 8d recipe)."""
    one = python_reference_sample(2, ROOTS[0], 1, 1)
    if one is None:
        return {"unavailable": "reference checkout not staged (scripts/stage_reference.sh -> baseline/_ref)"}
    if "unavailable" in one:
        return one
    many = python_reference_sample(2, ROOTS[0], threads, threads)
    leg = {"kind": "reference", "unit": "games/s",
           "impl": one["impl"], "call": "run_tournament.run_tournament(config=TournamentConfig(n_players=2, "
           "deterministic_batch_size=1), strategies=<5,160 grid>, global_seed=42, n_jobs=J, collect_metrics=True, "
           "row_output_directory=None)  [simulation/run_tournament.py:1050]",
           "n_jobs_1": {"value": one["games_per_s"], "cores": 1, "games": one["games"], "seconds": one["seconds"]}}
    if many and "unavailable" not in many:
        leg["n_jobs_all"] = {"value": many["games_per_s"], "cores": many["n_jobs"], "games": many["games"],
                             "seconds": many["seconds"], "wins_total": many["wins_total"]}
        leg["value"], leg["cores"] = many["games_per_s"], many["n_jobs"]
        leg["sample"] = (f"full grid root {ROOTS[0]} k=2: shuffles 0..{many['shuffles'] - 1} ({many['games']} games, "
                         f"{many['seconds']:.1f} s) on {many['n_jobs']} worker processes; 1 worker: shuffle 0 "
                         f"({one['games']} games, {one['seconds']:.1f} s)")
    return leg


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    table = full_grid_table()
    threads = os.cpu_count() or 1
    per_step = max(2.0, min(20.0, 120.0 / max(args.steps + args.warmup, 1)))
    for _ in range(args.warmup):
        cpu_sample(table, per_step / 4, threads)
    results = [cpu_sample(table, per_step, threads) for _ in range(args.steps)]
    games = sum(r["games"] for r in results)
    secs = sum(r["seconds"] for r in results)
    value = games / secs
    base = {k_: results[-1][k_] for k_ in ("unit", "cores", "kind", "sample")}
    pyref = python_reference_leg(threads) if args.ref_shuffles != 0 else {"unavailable": "skipped (--ref-shuffles 0)"}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "games/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * secs / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32/u64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, **base},
        "e2e": {"value": value, "unit": "games/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "cpu_baseline_reference": pyref,
        "note": ("value / cpu_baseline = C restatement of the reference (oracle/, pthreads over shuffles, "
                 "all host cores), bounded sample per step: the HARDER baseline.  cpu_baseline_reference "
                 "= the unmodified Python reference's own process-pool path timed on the same cores "
                 "(one shuffle per core, one sample: a shuffle costs ~11 core-seconds in Python, so "
                 "K+W such samples would not fit the arm's time budget)"),
    }
    print(json.dumps(line))


def workload_config(n_gpus: int) -> dict:
    return {
        "workload": ("configs/default_config.yaml as BASELINE.json describes it: full 5,160-strategy "
                     "grid, n_players_list [2,4], seeds [42,43]; one step = one root's k=2 and k=4 "
                     "cells at 4,300 shuffles (16,641,000 games) per GPU"),
        "n_strategies": N_STRATEGIES, "k": list(CELLS_K), "roots": list(ROOTS),
        "shuffles_per_cell_per_gpu": SHUFFLES, "shuffles_per_batch": SHUFFLES_PER_BATCH,
        "games_per_step_per_gpu": sum(SHUFFLES * (N_STRATEGIES // k) for k in CELLS_K),
        "pipeline": ("one fb_play_tournament_cells call per step: [k=2 cell, k=4 cell] + the next step's k=2 "
                     "cell as look-ahead (prepared, not played)"),
        "sharding": (f"rank r plays shuffles [r*{SHUFFLES},(r+1)*{SHUFFLES}) of each cell; "
                     "one int64 all-reduce of the tally tensor per cell" if n_gpus > 1
                     else "single GPU"),
        "l2": ("no flush needed: every step re-seeds 1.77 GB of seat records per cell (80 B x 22.2 M "
               "seats), 14x the 126 MB L2, and consecutive steps alternate roots; the kernels are "
               "not memory bound"),
    }


def strong_scaling_leg(play_cells, table, rank: int, world: int, reps: int, timer,
                       cells=None, n_strategies: int = N_STRATEGIES, batch: int = SHUFFLES_PER_BATCH) -> dict:
    """Strong scaling of a FIXED workload: one mega-config root (configs/farkle_mega_config.yaml: full
    grid, k in {2,3,4,5,6,8,10,12}, 4,300 shuffles each = 39,013,900 games).  The cells are dealt to the
    ranks along one line by estimated cost (run_tournament.plan_cells), every rank plays its segments
    through the pipelined cell list, ONE all-reduce merges the stacked tallies.  Rank 0 also plays
    the whole root alone -- no collective, the other ranks wait at a barrier -- so that both times
    come from this run, and the two tally sets are compared.  `timer(fn, reps) -> (ms, result)` times
    on the device (CUDA events); tests pass a wall-clock one and CPU stand-ins for the launches."""
    import torch
    import torch.distributed as dist

    from farkle_ii_b200 import run_tournament as frt

    mega = cells if cells is not None else [(MEGA_ROOT, k, SHUFFLES) for k in MEGA_K]
    multi = world > 1

    def run(**kw):
        return frt.run_cells(mega, table, batch_size=batch, play_cells=play_cells, **kw)

    n_ms, (t_all, tot_all) = timer(lambda: run(rank=rank, world=world), reps)
    if multi:
        t = torch.tensor([n_ms], dtype=torch.float64, device=t_all.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        n_ms = float(t.item())
        dist.barrier()
    one_ms, same = n_ms, True
    if multi and rank == 0:
        alone = frt.plan_cells(mega, n_strategies, 1, batch_size=batch)
        one_ms, (t_one, tot_one) = timer(lambda: run(rank=0, world=1, plan=alone), max(1, reps - 1))
        same = bool(torch.equal(t_one, t_all) and torch.equal(tot_one, tot_all))
    if multi:
        dist.barrier()
    games = sum(s * (n_strategies // k) for _, k, s in mega)
    out = {"workload": f"mega-config root {mega[0][0]}: full grid, k in {[k for _, k, _ in mega]}, "
                       f"{mega[0][2]} shuffles each ({games} games), fixed total work",
           "n_gpus": world, "n1_ms": one_ms, "nN_ms": n_ms, "speedup": one_ms / n_ms,
           "efficiency": one_ms / n_ms / world, "games_per_s": games / (n_ms * 1e-3),
           "identical_tallies_n1_vs_nN": same, "games_attempted": int(tot_all[:, 0].sum().item()),
           "reps": reps,
           "plan": [[(sg.k, sg.shuffle0, sg.n_shuffles) for sg in segs]
                    for segs in frt.plan_cells(mega, n_strategies, world, batch_size=batch)],
           "timing": "CUDA events on every rank around all launches + the all-reduce, max over ranks; "
                     "n1_ms: rank 0 plays the whole root alone in the same run"}
    assert out["games_attempted"] == games
    return out


# ----------------------------------------------------------------------------- GPU arm
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--shuffles", type=int, default=SHUFFLES, help=argparse.SUPPRESS)
    ap.add_argument("--unpipelined", action="store_true",
                    help="one fb_play_tournament call per cell instead of the pipelined cell list")
    ap.add_argument("--parquet-batches", type=int, default=20,
                    help="deterministic batches of the k=2 cell written to Parquet by the e2e_parquet leg; 0 skips it")
    ap.add_argument("--strong-reps", type=int, default=3,
                    help="timed repetitions of the strong-scaling leg (mega root over the ranks); 0 skips it")
    ap.add_argument("--ref-shuffles", type=int, default=-1,
                    help="0 skips the Python-reference CPU leg (cpu_baseline_reference)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    from farkle_ii_b200.device import get_engine
    from farkle_ii_b200.layout import TALLY_WIDTH, TOTALS_WIDTH

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # Libraries (NCCL prints its version banner) write to fd 1: keep the real stdout for the one
    # JSON line and send everything else to stderr.
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: farkle_ii_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        from datetime import timedelta

        # a collective that does not complete in 90 s is a bug, not a slow run (NCCL's default of
        # ten minutes would burn the box)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=timedelta(seconds=90))
    eng = get_engine(local)
    table_host = full_grid_table()
    table_dev = eng.to_device(table_host)
    n_sh = args.shuffles
    shuffle0 = rank * n_sh

    tallies = {k: torch.zeros((1, N_STRATEGIES, TALLY_WIDTH), dtype=torch.int64, device=eng.device)
               for k in CELLS_K}
    totals = {k: torch.zeros(TOTALS_WIDTH, dtype=torch.int64, device=eng.device) for k in CELLS_K}

    def step(i: int) -> None:
        # One call per step with the step's cells (fb_play_tournament_cells): the permutations and
        # seat seeding of a cell run under the tail / finish / tally passes of the cell before it.
        # The first cell of the NEXT step is passed as look-ahead, as a runner walking its list of
        # (root, k) cells does; every step therefore still prepares and plays two cells.
        root, nxt = ROOTS[i % len(ROOTS)], ROOTS[(i + 1) % len(ROOTS)]
        for k in CELLS_K:
            tallies[k].zero_()
            totals[k].zero_()
        if args.unpipelined:
            for k in CELLS_K:
                eng.play_tournament(root, k, shuffle0, n_sh, table_dev, tallies=tallies[k], totals=totals[k])
        else:
            eng.play_cells([(root, k, shuffle0, n_sh, tallies[k], totals[k]) for k in CELLS_K], table_dev,
                           ahead=(nxt, CELLS_K[0], shuffle0, n_sh))
        if world > 1:
            for k in CELLS_K:
                dist.all_reduce(tallies[k])

    def sync_all() -> None:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    sync_all()
    launches0 = eng.kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        sync_all()
        ev0.record()
        for i in range(args.steps):
            step(i)
        ev1.record()
        sync_all()
    ms = ev0.elapsed_time(ev1)
    launches = eng.kernel_launch_count() - launches0
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=eng.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    games_per_step_rank = sum(n_sh * (N_STRATEGIES // k) for k in CELLS_K)
    total_games = games_per_step_rank * world * args.steps
    value = total_games / (ms * 1e-3)

    # ---- dominant kernel: its launches INSIDE the timed region, timed by CUDA events the library
    # records on the launching stream around every play_kernel (newest first: per step k=4, k=2)
    hist = eng.play_kernel_ms_history(len(CELLS_K) * args.steps)
    kern = {}
    for j, k in enumerate(reversed(CELLS_K)):
        ms_k = hist[j::len(CELLS_K)]
        kern[k] = {"ms": float(np.mean(ms_k)), "launches_timed": len(ms_k)}
    for k in CELLS_K:  # the work counters of one launch per cell kind (deterministic per root)
        totals[k].zero_()
        tallies[k].zero_()
        eng.play_tournament(ROOTS[(args.steps - 1) % 2], k, shuffle0, n_sh, table_dev, tallies=tallies[k],
                            totals=totals[k])
        tot = totals[k].cpu().numpy()
        model_k = sass_model()["k2" if k == 2 else "generic"]
        alg, exe = lane_ops(tot, model_k)
        kern[k].update({"totals": tot, "games": int(tot[0]), "ops": alg, "ops_executed": exe,
                        "ops_survey": lane_ops(tot, SURVEY_MODEL)[0], "model": model_k})
    # e2e through the host-buffer C-ABI call
    out_t = {k: np.empty((1, N_STRATEGIES, TALLY_WIDTH), dtype=np.int64) for k in CELLS_K}
    pin = torch.from_numpy(table_host.view(np.uint8).copy()).pin_memory()
    table_pinned = pin.numpy().view(table_host.dtype)

    def e2e_step(i: int) -> None:
        root = ROOTS[i % len(ROOTS)]
        for k in CELLS_K:
            eng.run_tournament_host(root, k, shuffle0, n_sh, table_pinned, out_tallies=out_t[k])
            if world > 1:
                t = torch.from_numpy(out_t[k]).to(eng.device)
                dist.all_reduce(t)
                out_t[k][...] = t.cpu().numpy()

    e2e_step(0)
    sync_all()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 4))
    for i in range(e2e_steps):
        e2e_step(i)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=eng.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = games_per_step_rank * world * e2e_steps / e2e_s

    # e2e with the per-game rows as well: every game's compact row lands in pinned host memory
    # (the ingest hand-off), copied under the kernels of the next chunk of shuffles
    from farkle_ii_b200.layout import row_dtype

    rows_pin = {k: torch.empty(n_sh * (N_STRATEGIES // k) * row_dtype(k).itemsize,
                               dtype=torch.uint8).pin_memory() for k in CELLS_K}
    rows_host = {k: rows_pin[k].numpy().view(row_dtype(k)) for k in CELLS_K}

    def rows_step(i: int) -> None:
        root = ROOTS[i % len(ROOTS)]
        for k in CELLS_K:
            eng.run_tournament_host(root, k, shuffle0, n_sh, table_pinned, out_tallies=out_t[k],
                                    want_rows=True, out_rows=rows_host[k])

    rows_step(0)
    sync_all()
    t0 = time.perf_counter()
    for i in range(2):
        rows_step(i)
    torch.cuda.synchronize()
    rows_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([rows_s], dtype=torch.float64, device=eng.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        rows_s = float(t.item())
    rows_value = games_per_step_rank * world * 2 / rows_s
    rows_d2h = sum(rows_pin[k].numel() for k in CELLS_K)
    h2d = len(CELLS_K) * table_host.nbytes
    d2h = len(CELLS_K) * (N_STRATEGIES * TALLY_WIDTH * 8 + TOTALS_WIDTH * 8)

    # ---- strong scaling of a FIXED workload (see strong_scaling_leg)
    from farkle_ii_b200 import run_tournament as frt

    strong = None
    if args.strong_reps > 0:
        def cuda_timer(fn, reps):
            out = fn()                                   # warm-up
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                out = fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps, out

        sync_all()
        strong = strong_scaling_leg(lambda segments, table: eng.play_cells(segments, table), table_dev,
                                    rank, world, args.strong_reps, cuda_timer)

    # ---- rows mode all the way to disk (rank 0, N = 1 only): the Python surface's run_tournament()
    # with a row directory -- games -> compact rows -> D2H -> Arrow -> one Parquet shard + manifest
    # line per shuffle (the reference's on-disk contract, run_tournament.py:530-558) on a bounded
    # sample of the k=2 cell.  The host encode, not the GPU, sets this rate.
    e2e_parquet = None
    if world == 1 and args.parquet_batches > 0:
        import shutil
        import tempfile

        from farkle_ii_b200.strategies import generate_strategy_grid

        strategies = generate_strategy_grid()[0]
        n_pq = args.parquet_batches * SHUFFLES_PER_BATCH
        out = Path(tempfile.mkdtemp(prefix="fb_rows_"))
        try:
            cfg = frt.TournamentConfig(n_players=2, num_shuffles=n_pq, deterministic_batch_size=SHUFFLES_PER_BATCH)
            frt.run_tournament(config=frt.TournamentConfig(n_players=2, num_shuffles=SHUFFLES_PER_BATCH,
                                                           deterministic_batch_size=SHUFFLES_PER_BATCH),
                               global_seed=ROOTS[0], checkpoint_path=out / "warm" / "c.pkl",
                               row_output_directory=out / "warm" / "rows", num_shuffles=SHUFFLES_PER_BATCH,
                               strategies=strategies, resume=False)                      # warm-up: imports, pools
            for key in frt.IO_STATS:
                frt.IO_STATS[key] = 0
            t0 = time.perf_counter()
            frt.run_tournament(config=cfg, global_seed=ROOTS[0], checkpoint_path=out / "run" / "2p_checkpoint.pkl",
                               row_output_directory=out / "run" / "rows", num_shuffles=n_pq, strategies=strategies,
                               resume=False)
            dt = time.perf_counter() - t0
            shards = list((out / "run" / "rows").glob("rows_*.parquet"))
            games_pq = n_pq * (N_STRATEGIES // 2)
            stats = dict(frt.IO_STATS)
            threads = min(32, os.cpu_count() or 1)
            e2e_parquet = {
                "value": games_pq / dt, "unit": "games/s", "games": games_pq, "seconds": dt,
                "shards": len(shards), "bytes_on_disk": sum(p.stat().st_size for p in shards),
                "writer_threads": threads,
                "launch_wall_s": stats["launch_wall_s"], "arrow_build_cpu_s": stats["arrow_build_cpu_s"],
                "parquet_write_cpu_s": stats["parquet_write_cpu_s"],
                "bottleneck": max((("Parquet encode + fsync + rename", stats["parquet_write_cpu_s"] / threads),
                                   ("Arrow build", stats["arrow_build_cpu_s"] / threads),
                                   ("GPU launches + D2H", stats["launch_wall_s"])), key=lambda kv: kv[1])[0],
                "sample": f"k=2 full grid root {ROOTS[0]}, shuffles 0..{n_pq - 1}: one shard per shuffle",
                "call": "farkle_ii_b200.run_tournament.run_tournament(row_output_directory=...) -> "
                        "rows_{root}_{k}p_{shuffle:012d}.parquet + manifest.jsonl (temp -> fsync -> rename)"}
            assert len(shards) == n_pq
        finally:
            shutil.rmtree(out, ignore_errors=True)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of play_kernel (k=2 cell: the dominant launch of the step) ----
    peak_measured = eng.measure_issue_peak()
    clk = clocks.summary()
    dom = max(CELLS_K, key=lambda k_: kern[k_]["ms"])
    achieved = kern[dom]["ops"] / (kern[dom]["ms"] * 1e-3)
    tot = kern[dom]["totals"]
    # algorithmic HBM bytes of play_kernel per game: every 80-byte seat record is read once and
    # its 32 bytes of counters written back once, plus the 4-byte game header
    row_bytes_alg = kern[dom]["games"] * (dom * (80 + 32) + 4)
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    nominal_peak = eng.sm_count * 4 * 32 * (clk.get("sm_max_mhz") or 1965) * 1e6
    traffic = None
    tpath = ROOT / "profiles" / "play_kernel_traffic.json"
    if tpath.exists():
        try:
            traffic = json.loads(tpath.read_text()).get(f"k{dom}_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {
        "bound": "issue", "kernel": f"play_kernel (k={dom} cell, {kern[dom]['games']} games/launch)",
        "achieved": achieved / 1e12, "peak": peak_measured / 1e12, "unit": "Tlaneop/s",
        "frac": achieved / peak_measured,
        "frac_model": achieved / peak_measured,
        "frac_executed": kern[dom]["ops_executed"] / (kern[dom]["ms"] * 1e-3) / peak_measured,
        "frac_survey_constants": kern[dom]["ops_survey"] / (kern[dom]["ms"] * 1e-3) / peak_measured,
        "frac_by_k": {str(k_): {"model": kern[k_]["ops"] / (kern[k_]["ms"] * 1e-3) / peak_measured,
                                "executed": kern[k_]["ops_executed"] / (kern[k_]["ms"] * 1e-3) / peak_measured}
                      for k_ in CELLS_K},
        "peak_source": ("measured: fb_measure_issue_peak = best of four register-only integer-chain "
                        "probes at 1,024 threads/SM (the LOP3 + IMAD.IADD one issues at ~0.98 IPC: "
                        "profiles/r02_issue_peak.md); MEASURED_PEAKS.json holds no integer peak"),
        "peak_variants": {str(v): eng.measure_issue_peak_variant(v) / 1e12 for v in range(5)},
        "peak_nominal": nominal_peak / 1e12,
        "model": (f"algorithmic lane-instructions = {kern[dom]['model']['W']}*rng_words + "
                  f"{kern[dom]['model']['D']}*dice + {kern[dom]['model']['R']}*rolls: SURVEY.md §8d's formula "
                  "with SASS-exact constants of this kernel (scripts/sass_model.py on the ncu capture under "
                  "profiles/; SURVEY's estimates were 28/8/90 -> frac_survey_constants); words, dice and rolls "
                  "are returned by the kernel.  frac_executed charges everything the kernel executes per roll "
                  f"({kern[dom]['model'].get('E')} lane-instructions: the surplus words of the face-queue top-up "
                  "included) = issue-active x lane efficiency"),
        "per_game": {"rolls": float(tot[3] / tot[0]), "dice": float(tot[4] / tot[0]),
                     "rng_words": float(tot[5] / tot[0]),
                     "lane_ops": float(kern[dom]["ops"] / tot[0])},
        "kernel_ms": kern[dom]["ms"],
        "kernel_launches_timed": kern[dom]["launches_timed"],
        "kernel_timing": "CUDA events around every play_kernel launch of the timed region, on its stream",
        "kernel_ms_by_k": {str(k_): kern[k_]["ms"] for k_ in CELLS_K},
        "traffic": traffic,
        "hbm": {"bound": "hbm", "algorithmic_bytes_per_launch": row_bytes_alg,
                "achieved": row_bytes_alg / (kern[dom]["ms"] * 1e-3) / 1e9, "peak": hbm_peak,
                "unit": "GB/s", "frac": row_bytes_alg / (kern[dom]["ms"] * 1e-3) / 1e9 / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"},
    }
    cpu = cpu_sample(table_host, args.cpu_seconds, os.cpu_count() or 1) if world == 1 else None
    if cpu:
        cpu.pop("seconds", None)
        cpu.pop("games", None)
    pyref = None
    if world == 1 and args.ref_shuffles != 0:
        pyref = python_reference_leg(os.cpu_count() or 1)

    # ---- parity of what was just timed: one deterministic batch per cell against the oracle, the
    # step's own tallies against the path's invariants, and (when the Python reference ran) the
    # games it completed on its shuffles against the GPU's count for the same shuffles
    import oracle

    oracle.build()
    slot = 37
    parity = {"slot": slot, "equal": True, "cells": []}
    for k in CELLS_K:
        root = ROOTS[(args.steps - 1) % 2]
        s0 = shuffle0 + slot * SHUFFLES_PER_BATCH
        got = eng.play_tournament(root, k, s0, SHUFFLES_PER_BATCH, table_dev)
        want_t, want_tot, _ = oracle.play_tournament(root, k, s0, SHUFFLES_PER_BATCH, table_host,
                                                     n_threads=os.cpu_count() or 1)
        same = bool(np.array_equal(got.tallies.cpu().numpy().reshape(want_t.shape), want_t) and
                    np.array_equal(got.totals.cpu().numpy(), want_tot))
        # invariants of the full timed cell (tallies[k] / totals[k] hold its last launch)
        t_full, tot_full = tallies[k].cpu().numpy()[0], totals[k].cpu().numpy()
        inv = bool(t_full[:, 0].sum() == tot_full[1] and t_full[:, 1].sum() == k * tot_full[0] and
                   t_full[:, 2].sum() + t_full[:, 3].sum() == k * tot_full[0] and
                   tot_full[0] == n_sh * (N_STRATEGIES // k) and tot_full[7] == 0)
        parity["cells"].append({"root": root, "k": k, "shuffles": [int(s0), int(s0 + SHUFFLES_PER_BATCH)],
                                "games": int(want_tot[0]), "batch_equals_oracle": same,
                                "full_cell_invariants": inv})
        parity["equal"] = parity["equal"] and same and inv
    if pyref and "n_jobs_all" in pyref:
        n_ref = pyref["n_jobs_all"]["cores"]
        got = eng.play_tournament(ROOTS[0], 2, 0, n_ref, table_dev)
        tot = got.totals.cpu().numpy()
        ok = bool(int(tot[0]) == pyref["n_jobs_all"]["games"] and int(tot[1]) == pyref["n_jobs_all"]["wins_total"])
        parity["python_reference"] = {"root": ROOTS[0], "k": 2, "shuffles": [0, n_ref], "games": int(tot[0]),
                                      "games_completed_gpu": int(tot[1]),
                                      "games_completed_reference": pyref["n_jobs_all"]["wins_total"], "equal": ok}
        parity["equal"] = parity["equal"] and ok
    parity["checker"] = "oracle/farkle_oracle.c (pinned to the reference by tests/test_oracle_golden.py)"
    line = {
        "metric": METRIC, "value": value, "unit": "games/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32/u64", "data": "synthetic",
        "config": workload_config(world),
        "e2e": {"value": e2e_value, "unit": "games/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "call": "Engine.run_tournament_host -> fb_run_tournament_host (host buffers)"},
        "e2e_rows": {"value": rows_value, "unit": "games/s", "h2d_bytes_per_step": h2d,
                     "d2h_bytes_per_step": d2h + rows_d2h, "steps": 2,
                     "call": "fb_run_tournament_host with rows_host: tallies + one compact row per game "
                             "into pinned host memory, D2H overlapped with the next chunk's kernels"},
        "e2e_parquet": e2e_parquet,
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": roofline,
        "roofline_hbm": {**roofline["hbm"], "traffic": traffic,
                         "note": "same kernel against the HBM roofline: it is not the binding one"},
        "cpu_baseline": cpu,
        "cpu_baseline_reference": pyref,
        "parity_check": parity,
        "strong": strong,
        "published_reference": {"games_per_s_1_worker": 279.1, "games_per_s_12_workers": 1142.9,
                                "hardware": "Ryzen 7 3700X, fast grid k=2 (BASELINE.md §1)"},
    }
    json_out.write(json.dumps(line) + "\n")
    json_out.flush()
    if world > 1:
        dist.destroy_process_group()
    if not parity["equal"]:
        raise SystemExit("bench.py: parity check failed: " + json.dumps(parity))
    if strong and not strong["identical_tallies_n1_vs_nN"]:
        raise SystemExit("bench.py: the N-rank tallies of the strong-scaling leg differ from the 1-rank ones")


if __name__ == "__main__":
    main()
