"""Synthetic workloads of the BASELINE.json configs, generated on the host from seeds.

Every generator returns plain numpy arrays in the layout the C ABI takes (folded draws ``(S, J)``), so bench.py,
the tests and ``__graft_entry__.smoke()`` all run on identical inputs.  Parameter draws follow the live Stan
model's priors and transforms (pyipn/stan_models/rff_omega_ipn.stan:100-124, 152-177).
"""
from __future__ import annotations

import math

import numpy as np

from .params import fold_draws

SEED = 20261019


def draw_prior(rng, k, max_range):
    """One parameter draw: (omega (2,k), beta1 (k,), beta2 (k,), scale (2,)) from the model's priors."""
    while True:
        raw = np.sort(rng.normal(1.0, 1.0, size=2))  # positive_ordered[2] ~ N(1,1)
        if raw[0] > 0:
            break
    while True:
        r1 = rng.lognormal(0.0, 0.2)  # range1_raw in (0,1)
        if 0 < r1 < 1:
            break
    while True:
        r2 = rng.lognormal(math.log(1e-2), 0.2)  # range2_raw in (0, range1_raw)
        if 0 < r2 < r1:
            break
    scale = raw / math.sqrt(k)
    bw = 1.0 / (np.array([r1, r2]) * max_range)
    omega = rng.normal(size=(2, k)) * bw[:, None]
    return omega, rng.normal(size=k), rng.normal(size=k), scale


def _intensity(time, omega, beta1, beta2, scale):
    """float64 numpy evaluation used ONLY to synthesise the observed counts (data generation, not the product path)."""
    f1 = np.outer(time, omega[0])
    f2 = np.outer(time, omega[1])
    return np.exp((scale[0] * np.cos(f1) + scale[1] * np.cos(f2)) @ beta1 + (scale[0] * np.sin(f1) + scale[1] * np.sin(f2)) @ beta2)


def delay_grid_workload(T=100_000, G=4096, S=1, k=50, dt_bin=1e-3, true_delay=0.1234, amp=0.8, boost=400.0, bkg=500.0,
                        seed=SEED, jitter=True):
    """Configs 2 (S = 1) and 3 (S = 4000 posterior-like draws around the truth) of BASELINE.json.

    Detector 1 of a 2-detector pair: bins of ``dt_bin`` s, a bright RFF burst (source rate = boost * amp * f) over a
    flat ``bkg`` cps background, delayed by ``true_delay``; trial delays ``linspace(-0.5, 0.5, G)``.
    """
    rng = np.random.default_rng(seed)
    edges = dt_bin * np.arange(T + 1)
    mid = edges[:-1] + 0.5 * dt_bin
    exposure = np.full(T, dt_bin)
    omega0, b1, b2, sc = draw_prior(rng, k, mid[-1] - mid[0])
    A_true = boost * amp
    # counts are generated in chunks to keep the (T, k) temporaries small
    counts = np.empty(T, dtype=np.int64)
    for lo in range(0, T, 20000):
        hi = min(T, lo + 20000)
        f = _intensity(mid[lo:hi] - true_delay, omega0, b1, b2, sc)
        counts[lo:hi] = rng.poisson(exposure[lo:hi] * (A_true * f + bkg))
    # posterior-like draws: truth + jitter (first draw is the truth itself)
    omega1 = np.repeat(omega0[0][:, None], S, axis=1)
    omega2 = np.repeat(omega0[1][:, None], S, axis=1)
    beta1 = np.repeat(b1[:, None], S, axis=1)
    beta2 = np.repeat(b2[:, None], S, axis=1)
    scale = np.repeat(sc[:, None], S, axis=1)
    amps = np.full(S, A_true)
    bkgs = np.full(S, bkg)
    if jitter and S > 1:
        bw = np.array([np.std(omega0[0]), np.std(omega0[1])])
        beta1[:, 1:] += rng.normal(0, 0.05, size=(k, S - 1))
        beta2[:, 1:] += rng.normal(0, 0.05, size=(k, S - 1))
        omega1[:, 1:] += rng.normal(0, 0.01, size=(k, S - 1)) * bw[0]
        omega2[:, 1:] += rng.normal(0, 0.01, size=(k, S - 1)) * bw[1]
        scale[:, 1:] *= np.exp(rng.normal(0, 0.02, size=(2, S - 1)))
        amps[1:] *= np.exp(rng.normal(0, 0.02, size=S - 1))
        bkgs[1:] *= np.exp(rng.normal(0, 0.02, size=S - 1))
    omega, a, b = fold_draws(omega1, omega2, beta1, beta2, np.ones(S), scale)
    delays = np.linspace(-0.5, 0.5, G)
    return dict(time=mid, exposure=exposure, counts=counts, omega=omega, a=a, b=b, amp=amps, bkg=bkgs, delays=delays,
                true_delay=true_delay, ref=dict(omega1=omega1, omega2=omega2, beta1=beta1, beta2=beta2, scale=scale))


def healpix_ring_vectors(nside):
    """Unit vectors of all 12 nside^2 HEALPix pixel centres, RING ordering (vectorised; Gorski et al. 2005)."""
    npix = 12 * nside * nside
    ncap = 2 * nside * (nside - 1)
    p = np.arange(npix, dtype=np.int64)
    z = np.empty(npix)
    phi = np.empty(npix)
    # north cap
    m = p < ncap
    pp = p[m]
    iring = (1 + np.floor(np.sqrt(1 + 2 * pp.astype(np.float64))).astype(np.int64)) // 2
    # guard against float sqrt at ring boundaries
    iring = np.where(2 * iring * (iring - 1) > pp, iring - 1, iring)
    iring = np.where(2 * (iring + 1) * iring <= pp, iring + 1, iring)
    iphi = pp + 1 - 2 * iring * (iring - 1)
    z[m] = 1.0 - iring * iring * 4.0 / npix
    phi[m] = (iphi - 0.5) * np.pi / (2.0 * iring)
    # equatorial belt
    m = (p >= ncap) & (p < npix - ncap)
    ip = p[m] - ncap
    iring = ip // (4 * nside) + nside
    iphi = ip % (4 * nside) + 1
    fodd = np.where(((iring + nside) & 1) == 1, 1.0, 0.5)
    z[m] = (2 * nside - iring) * 2.0 / (3.0 * nside)
    phi[m] = (iphi - fodd) * np.pi / (2.0 * nside)
    # south cap
    m = p >= npix - ncap
    ip = npix - p[m]
    iring = (1 + np.floor(np.sqrt(2 * ip.astype(np.float64) - 1)).astype(np.int64)) // 2
    iring = np.where(2 * iring * (iring - 1) >= ip, iring - 1, iring)
    iring = np.where(2 * (iring + 1) * iring < ip, iring + 1, iring)
    iphi = 4 * iring + 1 - (ip - 2 * iring * (iring - 1))
    z[m] = -1.0 + iring * iring * 4.0 / npix
    phi[m] = (iphi - 0.5) * np.pi / (2.0 * iring)
    st = np.sqrt(np.maximum(0.0, (1.0 - z) * (1.0 + z)))
    return np.stack([st * np.cos(phi), st * np.sin(phi), z], axis=1)


def sky_grid_workload(D=8, T=100_000, S=1, nside=64, k=50, dt_bin=1e-3, boost=400.0, bkg=500.0, seed=SEED, radii=None):
    """Config 4: D detectors at synthetic positions (km), each with T bins covering its own arrival window."""
    rng = np.random.default_rng(seed + 4)
    if radii is None:
        radii = np.array([7.0e3, 4.2e4, 3.8e5, 1.5e6, 1.5e6, 5.0e7, 1.5e8, 2.3e8])[:D]
    dirs = rng.normal(size=(D, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    sc_pos = dirs * np.asarray(radii)[:, None]
    g_true = rng.normal(size=3)
    g_true /= np.linalg.norm(g_true)
    C = 299792.46
    true_delays = (sc_pos[0][None, :] - sc_pos) @ g_true / C  # (D,), detector 0 -> 0
    span = T * dt_bin
    omega0, b1, b2, sc = draw_prior(rng, k, span)
    time = np.zeros((D, T))
    exposure = np.full((D, T), dt_bin)
    counts = np.zeros((D, T), dtype=np.int64)
    amps = np.ones((D, S))
    amps[1:] = np.exp(rng.normal(0, 0.3, size=(D - 1, 1)))
    bkgs = np.full((D, S), bkg) * np.exp(rng.normal(0, 0.05, size=(D, 1)))
    for d in range(D):
        # detector d records the window [delay_d, delay_d + span) in its own clock so that t - delay spans [0, span)
        mid = true_delays[d] + dt_bin * (np.arange(T) + 0.5)
        time[d] = mid
        for lo in range(0, T, 20000):
            hi = min(T, lo + 20000)
            f = _intensity(mid[lo:hi] - true_delays[d], omega0, b1, b2, sc)
            counts[d, lo:hi] = rng.poisson(exposure[d, lo:hi] * (boost * amps[d, 0] * f + bkgs[d, 0]))
    amps = amps * boost
    omega, a, b = fold_draws(np.repeat(omega0[0][:, None], S, 1), np.repeat(omega0[1][:, None], S, 1),
                             np.repeat(b1[:, None], S, 1), np.repeat(b2[:, None], S, 1), np.ones(S),
                             np.repeat(sc[:, None], S, 1))
    ghat = healpix_ring_vectors(nside)
    return dict(nbins=np.full(D, T, dtype=np.int64), time=time, exposure=exposure, counts=counts, sc_pos=sc_pos, ghat=ghat,
                omega=omega, a=a, b=b, amp=amps, bkg=bkgs, g_true=g_true, true_delays=true_delays)
