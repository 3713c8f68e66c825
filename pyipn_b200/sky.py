"""Sky post-processing: from time delays and per-pixel likelihoods to localisation products.

Host-side mirrors of the reference's small geometry helpers, on plain km vectors instead of astropy objects, plus the
grid-native credible regions computed by ``libipn_b200`` (``ipn_sky_credible``):

* ``grb_phi_theta``                 generated quantities of pyipn/stan_models/rff_omega_ipn.stan:229-230
* ``theta_from_time_delay``         pyipn/utils/timing.py:14-23
* ``calculate_distance_and_norm``   pyipn/utils/timing.py:26-59
* ``calculate_annulus`` / ``localize_GRB``   pyipn/universe.py:364-386, 670-697
* ``credible_regions``              replaces the KDE -> MOC route of pyipn/fit.py:193-206, 328-347
"""
from __future__ import annotations

from itertools import combinations

import numpy as np

from ._lib import default_context

# astropy.constants.c in km/s: the classical (annulus) helpers use it, while the Stan likelihood hard-codes 299792.46
SPEED_OF_LIGHT_KM_S = 299792.458


def grb_phi_theta(grb_xyz):
    """(phi, theta) = (atan2(y, x), -(acos(z) - pi/2)): longitude and latitude of a unit vector, radians."""
    g = np.asarray(grb_xyz, dtype=np.float64)
    return np.arctan2(g[..., 1], g[..., 0]), 0.5 * np.pi - np.arccos(g[..., 2])


def theta_from_time_delay(dt, distance):
    """Half-opening angle (rad) of the annulus for a delay ``dt`` (s) over a baseline ``distance`` (km).  The ratio is
    rounded to 15 decimals like the reference does, so that |c dt / d| marginally above 1 does not give NaN."""
    return np.arccos(np.around(SPEED_OF_LIGHT_KM_S * np.asarray(dt, dtype=np.float64) / np.asarray(distance, dtype=np.float64), 15))


def calculate_distance_and_norm(p1, p2):
    """Baseline from detector 1 to detector 2 (km vectors): ``(distance, unit vector, ra, dec)``, angles in radians."""
    d = np.asarray(p2, dtype=np.float64) - np.asarray(p1, dtype=np.float64)
    distance = float(np.linalg.norm(d))
    if not distance > 0.0:
        raise ValueError("the two detectors are at the same position")
    n = d / distance
    return distance, n, float(np.arctan2(n[1], n[0]) % (2.0 * np.pi)), float(np.arcsin(n[2]))


def calculate_annulus(sc_pos, T0, i, j):
    """``(unit vector i->j, (ra, dec), theta)`` of the annulus of detectors ``i`` and ``j`` with trigger times ``T0``."""
    distance, n, ra, dec = calculate_distance_and_norm(sc_pos[i], sc_pos[j])
    return n, np.array([ra, dec]), float(theta_from_time_delay(T0[i] - T0[j], distance))


def localize_GRB(sc_pos, T0):
    """Least-squares triangulation: the direction g minimising sum over detector pairs of (n_ij . g - cos theta_ij)^2.
    Returns the (not normalised) solution like the reference; divide by its norm for a unit vector."""
    sc_pos = np.asarray(sc_pos, dtype=np.float64)
    if sc_pos.ndim != 2 or sc_pos.shape[1] != 3 or sc_pos.shape[0] < 2:
        raise ValueError("sc_pos must be (D >= 2, 3)")
    rows, rhs = [], []
    for i, j in combinations(range(sc_pos.shape[0]), 2):
        n, _, theta = calculate_annulus(sc_pos, T0, i, j)
        rows.append(n)
        rhs.append(np.cos(theta))
    return np.linalg.lstsq(np.asarray(rows), np.asarray(rhs), rcond=None)[0]


def credible_regions(ll, levels=(0.68, 0.95), log_weight=None, pixel_area=None, ctx=None):
    """Credible regions of a per-pixel log-likelihood ``ll`` (P,) or (P, S) on an equal-area pixel grid.

    Returns a dict: ``prob`` (P,), ``level`` (P,), ``order`` (P,), ``log_norm``, ``best`` (arg-max pixel) and, per
    requested level c, ``masks[c]`` (pixels of the smallest highest-density set holding at least c) and ``areas[c]``
    (pixel count, or steradians when ``pixel_area`` is given; nside -> 4 pi / (12 nside^2))."""
    ctx = ctx or default_context()
    prob, level, order, log_norm = ctx.sky_credible(ll, log_weight)
    masks, areas = {}, {}
    for c in levels:
        if not 0.0 < c <= 1.0:
            raise ValueError("credible levels must be in (0, 1]")
        m = (level - prob) < c
        masks[c] = m
        areas[c] = float(m.sum()) * (1.0 if pixel_area is None else float(pixel_area))
    return dict(prob=prob, level=level, order=order, log_norm=log_norm, best=int(order[0]), masks=masks, areas=areas)
