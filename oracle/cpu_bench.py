"""Times the oracle port of the reference path on the host cores (bench.py's cpu_baseline / --impl reference).

TEST/BENCH INFRASTRUCTURE ONLY (see oracle/ipn_oracle.py header).  Each worker evaluates the direct
(no angle-addition shortcut) float64 composition the reference implies -- RFF_multiscale arithmetic
(pyipn/rff.py:19-29) on ``time - delay`` + Stan's poisson_lpmf (functions.stan:23-35) -- for a few delays of
one draw over all T bins, the same per-unit work the reference's numba path does (it has no parallel=True, so
more than one core means one process per core).
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np


_W = None


def _init(w):
    global _W
    _W = w


def _worker(args):
    from oracle import ipn_oracle as orc

    s, delay_idx = args
    w = _W
    t0 = time.perf_counter()
    out = [orc.loglik_folded(w["counts"], w["time"], w["exposure"], w["omega"][s], w["a"][s], w["b"][s], w["delays"][g],
                             w["bkg"][s], w["amp"][s]) for g in delay_idx]
    return time.perf_counter() - t0, out


def time_c_oracle(workload, pairs_per_core=32, cores=None):
    """The plain-C restatement (oracle/ipn_oracle.c: scalar libm calls in the reference's evaluation order, which is what
    numba makes of rff.py) on ``cores`` threads.  Returns dict(value=units/s, cores, sample, seconds)."""
    from oracle import c_oracle as corc

    cores = cores or os.cpu_count() or 1
    corc.set_threads(cores)
    T = len(workload["time"])
    G = len(workload["delays"])
    S = workload["omega"].shape[0]
    n_delays = max(1, cores * pairs_per_core // S)
    idx = (np.arange(n_delays) * 37) % G
    args = (workload["counts"], workload["time"], workload["exposure"], workload["omega"], workload["a"], workload["b"],
            workload["amp"], workload["bkg"])
    corc.loglik_delay_grid(*args, workload["delays"][idx[: max(1, cores // S)]])  # warm-up: page-in, thread start
    t0 = time.perf_counter()
    corc.loglik_delay_grid(*args, workload["delays"][idx])
    wall = time.perf_counter() - t0
    corc.set_threads(0)
    units = n_delays * S * T
    return dict(value=units / wall, cores=cores, seconds=wall,
                sample=f"{n_delays * S} (delay,draw) pairs x {T} bins on {cores} threads (C restatement of rff.py + Stan "
                       f"poisson_lpmf, float64, scalar libm)")


def time_oracle(workload, delays_per_task=8, tasks_per_core=1, cores=None, use_c=True):
    """Returns dict(value=units/s, cores, sample, seconds).  units = (delay, draw) pairs x bins.  The C restatement is
    timed when it is available (it runs at the speed of the reference's numba-compiled loops); the numpy port, one
    process per core, is the fall-back."""
    if use_c:
        try:
            return time_c_oracle(workload, cores=cores)
        except Exception as exc:  # no compiler and no prebuilt library: say so and time the numpy port
            print(f"[cpu_bench] C oracle unavailable ({exc}); timing the numpy port", flush=True)
    cores = cores or os.cpu_count() or 1
    T = len(workload["time"])
    G = len(workload["delays"])
    S = workload["omega"].shape[0]
    w = {k: workload[k] for k in ("counts", "time", "exposure", "omega", "a", "b", "delays", "bkg", "amp")}
    tasks = []
    for t in range(cores * tasks_per_core):
        s = t % S
        idx = [(t * delays_per_task + j) * 37 % G for j in range(delays_per_task)]
        tasks.append((s, idx))
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores, initializer=_init, initargs=(w,)) as pool:
        pool.map(_worker, tasks[:cores])  # warm-up: imports, page-in
        t0 = time.perf_counter()
        pool.map(_worker, tasks)
        wall = time.perf_counter() - t0
    units = len(tasks) * delays_per_task * T
    return dict(value=units / wall, cores=cores, seconds=wall,
                sample=f"{len(tasks) * delays_per_task} (delay,draw) pairs x {T} bins over {cores} processes "
                       f"(float64 numpy port of rff.py + Stan poisson_lpmf)")
