"""Host-side data generator mirroring ``pyipn/possion_gen.py`` (Norris/FRED pulse + thinning).

This is the simulator that produces test/bench INPUTS; it is not on the hot
path, so it stays on the host like the reference's numba version.  numba's ``np.random.seed/rand`` stream
equals ``numpy.random.RandomState`` for the same seed, so a run here reproduces the reference's arrival
times draw for draw (the reference's own generator no longer runs under numba 0.65).
"""
from __future__ import annotations

import numpy as np


def norris(x, K, t_start, t_rise, t_decay):
    """pyipn/possion_gen.py:9-33, vectorised over x."""
    x = np.asarray(x, dtype=np.float64)
    out = np.zeros_like(x)
    m = x > t_start
    d = x[m] - t_start
    out[m] = K * np.exp(2.0 * np.sqrt(t_rise / t_decay)) * np.exp((-t_rise / d) - d / t_decay)
    return out


def pulse(x, K, t_start, t_rise, t_decay):
    """pyipn/possion_gen.py:36-62,141-156: single pulse (scalars) or sum of pulses (arrays)."""
    if np.ndim(K) == 0:
        return norris(x, K, t_start, t_rise, t_decay)
    out = np.zeros_like(np.asarray(x, dtype=np.float64))
    for n in range(len(K)):
        out = out + norris(x, K[n], t_start[n], t_rise[n], t_decay[n])
    return out


def _thin(rs, tstart, tstop, fmax, rate_fn):
    """Shared thinning loop (possion_gen.py:196-204 / 232-240): step ~ Exp(fmax), accept w.p. rate/fmax.

    Consumes the RandomState stream in the reference's order (step draw, then test draw, per iteration)."""
    times = [np.array([tstart])]  # the reference appends tstart itself as the first "arrival"
    t = tstart
    while t < tstop:
        n = max(1024, int((tstop - t) * fmax * 1.1) + 64)
        u = rs.random_sample(2 * n).reshape(n, 2)
        cand = t + np.cumsum(-np.log(u[:, 0]) / fmax)
        # the loop body runs for every iteration that STARTS with time < tstop, i.e. up to and including the
        # first candidate >= tstop
        beyond = np.nonzero(cand >= tstop)[0]
        last = beyond[0] if beyond.size else n - 1
        cand = cand[: last + 1]
        keep = u[: last + 1, 1] <= rate_fn(cand) / fmax
        times.append(cand[keep])
        t = cand[-1]
        if beyond.size:
            break
    return np.concatenate(times)


def source_poisson_generator(tstart, tstop, K, p_start, t_rise, t_decay, seed=1234):
    """pyipn/possion_gen.py:159-206."""
    rs = np.random.RandomState(seed)
    grid = np.linspace(tstart, tstop + 1.0, 1000)
    fmax = pulse(grid, K, p_start, t_rise, t_decay).max()
    if not fmax > 0.0:
        return np.zeros(0)
    return _thin(rs, tstart, tstop, fmax, lambda t: pulse(t, K, p_start, t_rise, t_decay))


def background_poisson_generator(tstart, tstop, slope, intercept, seed=1234):
    """pyipn/possion_gen.py:209-242."""
    rs = np.random.RandomState(seed)
    grid = np.linspace(tstart, tstop + 1.0, 1000)
    fmax = (intercept + slope * grid).max()
    return _thin(rs, tstart, tstop, fmax, lambda t: intercept + slope * t)


def source_poisson_generator_gpu(tstart, tstop, K, p_start, t_rise, t_decay, seed=1234):
    """Device version of :func:`source_poisson_generator` (``ipn_thinning_events``): same point process, Philox stream."""
    from ._lib import default_context

    return default_context().thinning_events(tstart, tstop, K, p_start, t_rise, t_decay, 0.0, 0.0, seed)


def background_poisson_generator_gpu(tstart, tstop, slope, intercept, seed=1234):
    """Device version of :func:`background_poisson_generator`."""
    from ._lib import default_context

    return default_context().thinning_events(tstart, tstop, (), (), (), (), slope, intercept, seed)
