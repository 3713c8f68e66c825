"""The north-star evaluations: binned Poisson log-likelihood of the RFF intensity over a grid of
(delay | sky direction) x (posterior draw).

Likelihood definition: pyipn/stan_models/functions.stan:23-35 (per detector), 37-61 (detector sum);
sky -> delay: functions.stan:65-69, 164-169 with c = 299792.46 km/s; detector 0 is the unshifted reference
(rff_omega_ipn.stan:63-68, 129-134, 144-145).  All heavy lifting happens in libipn_b200 (CUDA, sm_100a).
"""
from __future__ import annotations

import numpy as np

from ._lib import default_context

C_KM_S = 299792.46  # functions.stan:65-69


def time_delay(grb_xyz, sc_pos_diff):
    """functions.stan:164-169 (host helper for single vectors; the sky grid computes this on the GPU)."""
    return np.dot(np.asarray(sc_pos_diff, dtype=np.float64), np.asarray(grb_xyz, dtype=np.float64)) * (1.0 / C_KM_S)


def sky_delays(ghat, sc_pos):
    """delays[p, d] = ghat_p . (sc_pos[0] - sc_pos[d]) / c  (host helper, float64)."""
    ghat = np.atleast_2d(np.asarray(ghat, dtype=np.float64))
    sc_pos = np.asarray(sc_pos, dtype=np.float64)
    return (ghat @ (sc_pos[0][None, :] - sc_pos).T) * (1.0 / C_KM_S)


def loglik_delay_grid(time, exposure, counts, omega, a, b, amp, bkg, delays, ctx=None):
    """ll[g, s] = sum_i poisson_lpmf(counts_i | exposure_i (amp_s f_s(time_i - delays_g) + bkg_s)); (G, S) float64."""
    ctx = ctx or default_context()
    return ctx.loglik_delay_grid(time, exposure, counts, omega, a, b, amp, bkg, delays)


def loglik_sky_grid(nbins, time, exposure, counts, sc_pos, ghat, omega, a, b, amp, bkg, ctx=None):
    """ll[p, s] = sum_d ll_d(delay[p, d]) with padded (D, Tmax) detector arrays; (P, S) float64."""
    ctx = ctx or default_context()
    return ctx.loglik_sky_grid(nbins, time, exposure, counts, sc_pos, ghat, omega, a, b, amp, bkg)


def argmax_with_gap(ll):
    """Per-draw argmax over the grid axis 0 and the best-minus-second-best gap in nats (hard part 6)."""
    ll = np.asarray(ll)
    idx = np.argmax(ll, axis=0)
    part = np.partition(ll, -2, axis=0) if ll.shape[0] > 1 else np.vstack([ll, ll])
    return idx, part[-1] - part[-2]


def marginal_over_draws(ll):
    """log mean_s exp(ll[g, s]) per grid point (stable log-sum-exp)."""
    ll = np.asarray(ll)
    m = ll.max(axis=1, keepdims=True)
    return (m + np.log(np.mean(np.exp(ll - m), axis=1, keepdims=True)))[:, 0]
