"""CPU oracle for the pyipn RFF-intensity -> binned-Poisson-log-likelihood hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pyipn_b200/`` may import this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs use it, and only as the checker (never as the thing measured as "ours" or shipped).

It is a float64 numpy restatement of the reference algorithm, function by function.  Citations are
``path:line`` relative to the reference checkout (grburgess/pyipn):

* ``rff`` / ``rff_multiscale``            <- pyipn/rff.py:5-16, 19-29
* ``expected_rate`` / ``_multiscale``     <- pyipn/fit.py:651-668, 671-691
* ``ppc_counts``                          <- pyipn/fit.py:694-705
* ``poisson_lpmf`` / ``loglik_bins``      <- pyipn/stan_models/functions.stan:23-35 (Stan math
                                             ``poisson_lpmf``: third-party, un-vendored, unpinned)
* ``total_loglik``                        <- pyipn/stan_models/functions.stan:37-61
* ``time_delay`` / ``C_KM_S``             <- pyipn/stan_models/functions.stan:65-69, 164-169;
                                             pyipn/stan_models/rff_omega_ipn.stan:63-68, 129-134, 144-145
* ``draw_prior``                          <- pyipn/stan_models/rff_omega_ipn.stan:100-124, 152-177
* ``bin_events`` / ``to_stan_data``       <- pyipn/lightcurve.py:24-47; pyipn/universe.py:441-529
* ``grb_angles``                          <- pyipn/stan_models/rff_omega_ipn.stan:227-233

Parity pinning (see DESIGN.md "Oracle"):
* The intensity functions are pinned against the *reference's own* ``pyipn/rff.py`` (imported by file
  path in the build container) through the committed fixtures in ``tests/golden/`` (generator:
  ``tests/golden/make_golden.py``).
* The draw loops, ``ppc_counts`` (MT19937 stream), ``bin_events`` and ``to_stan_data`` are pinned against the
  reference's own ``_expeced_rate*`` / ``ppc_generator`` / ``get_binned_light_curve`` / ``to_stan_data``, lifted out of
  ``fit.py`` / ``lightcurve.py`` / ``universe.py`` and run in the build container (``tests/golden/make_golden_path.py``
  -> ``path_golden.npz``, ``tests/test_path_golden.py``).
* The Poisson likelihood arithmetic lives in Stan math, which is not vendored by the reference and has
  no golden vector in the reference's tests  ->  **parity unpinned** at that boundary; the restatement
  follows Stan's documented ``poisson_lpmf`` (n log(mu) - mu - lgamma(n+1); mu=0,n=0 -> 0;
  mu=0,n>0 -> -inf) and is cross-checked against ``scipy.stats.poisson.logpmf`` in the tests.
* HEALPix RING ``pix2vec`` (used for the sky grid only) restates the published HEALPix algorithm
  (Gorski et al. 2005); the reference never builds a likelihood sky grid -> parity unpinned there too.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.special import gammaln

# pyipn/stan_models/functions.stan:65-69  (note: NOT 299792.458)
C_KM_S = 299792.46


# ----------------------------------------------------------------------------------------------
# intensity  (pyipn/rff.py)
# ----------------------------------------------------------------------------------------------
def rff(time, omega1, omega2, beta1, beta2, bw, scale):
    """pyipn/rff.py:5-16.  exp(scale * [(cos F1 + cos F2).b1 + (sin F1 + sin F2).b2]), F = bw * t (x) w."""
    time = np.asarray(time, dtype=np.float64)
    f1 = bw * np.outer(time, np.asarray(omega1, dtype=np.float64))
    f2 = bw * np.outer(time, np.asarray(omega2, dtype=np.float64))
    log_fhat = scale * (
        np.dot(np.cos(f1) + np.cos(f2), np.asarray(beta1, dtype=np.float64))
        + np.dot(np.sin(f1) + np.sin(f2), np.asarray(beta2, dtype=np.float64))
    )
    return np.exp(log_fhat)


def rff_multiscale(time, omega1, omega2, beta1, beta2, bw, scale1, scale2):
    """pyipn/rff.py:19-29.  exp((s1 cos F1 + s2 cos F2).b1 + (s1 sin F1 + s2 sin F2).b2)."""
    time = np.asarray(time, dtype=np.float64)
    f1 = bw * np.outer(time, np.asarray(omega1, dtype=np.float64))
    f2 = bw * np.outer(time, np.asarray(omega2, dtype=np.float64))
    log_fhat = np.dot(scale1 * np.cos(f1) + scale2 * np.cos(f2), np.asarray(beta1, dtype=np.float64)) + np.dot(
        scale1 * np.sin(f1) + scale2 * np.sin(f2), np.asarray(beta2, dtype=np.float64)
    )
    return np.exp(log_fhat)


def expected_rate(time, omega1, omega2, beta1, beta2, bw, scale, amplitude, dt, N=None):
    """pyipn/fit.py:651-668.  out[n] = amplitude[n] * RFF(time - dt[n], w1[:,n], w2[:,n], b1[:,n], b2[:,n], bw[n], scale[n])."""
    time = np.asarray(time, dtype=np.float64)
    N = len(amplitude) if N is None else N
    out = np.empty((N, len(time)))
    for n in range(N):
        out[n] = amplitude[n] * rff(
            time - dt[n], omega1[:, n], omega2[:, n], beta1[:, n], beta2[:, n], bw[n], scale[n]
        )
    return out


def expected_rate_multiscale(time, omega1, omega2, beta1, beta2, bw, scale, amplitude, dt, N=None):
    """pyipn/fit.py:671-691.  scale is (2, S): scale[0, n], scale[1, n]."""
    time = np.asarray(time, dtype=np.float64)
    N = len(amplitude) if N is None else N
    out = np.empty((N, len(time)))
    for n in range(N):
        out[n] = amplitude[n] * rff_multiscale(
            time - dt[n], omega1[:, n], omega2[:, n], beta1[:, n], beta2[:, n], bw[n], scale[0, n], scale[1, n]
        )
    return out


def ppc_counts(time, exposure, rates, bkg, seed):
    """pyipn/fit.py:694-705.  out[n, i] = Poisson((rates[n, i] + bkg[n]) * exposure[i]).

    numba's ``np.random.poisson`` uses the MT19937 legacy stream; ``numpy.random.RandomState`` with the
    same seed reproduces it draw for draw (hard part 8).  Row-major draw order.
    """
    rs = np.random.RandomState(seed)
    rates = np.asarray(rates, dtype=np.float64)
    N, T = rates.shape
    out = np.empty((N, T))
    for n in range(N):
        for i in range(T):
            out[n, i] = rs.poisson((rates[n, i] + bkg[n]) * exposure[i])
    return out


# ----------------------------------------------------------------------------------------------
# canonical folded form shared with the CUDA library (host-side algebra, exact in exact arithmetic)
#   log f(tau) = sum_j a_j cos(w_j tau) + b_j sin(w_j tau),  J = 2k
# ----------------------------------------------------------------------------------------------
def fold_params(omega1, omega2, beta1, beta2, bw, scale1, scale2):
    """Fold one draw (k,) x4 + scalars into (omega[J], a[J], b[J]) with J = 2k.

    single-scale RFF (rff.py:11-14) is scale1 == scale2 == scale.
    """
    omega = np.concatenate([bw * np.asarray(omega1, np.float64), bw * np.asarray(omega2, np.float64)])
    a = np.concatenate([scale1 * np.asarray(beta1, np.float64), scale2 * np.asarray(beta1, np.float64)])
    b = np.concatenate([scale1 * np.asarray(beta2, np.float64), scale2 * np.asarray(beta2, np.float64)])
    return omega, a, b


def log_intensity_folded(tau, omega, a, b):
    """sum_j a_j cos(w_j tau) + b_j sin(w_j tau) for a vector tau (float64)."""
    ph = np.outer(np.asarray(tau, np.float64), omega)
    return np.cos(ph) @ a + np.sin(ph) @ b


# ----------------------------------------------------------------------------------------------
# likelihood  (pyipn/stan_models/functions.stan)
# ----------------------------------------------------------------------------------------------
def poisson_lpmf(n, mu):
    """Sum over bins of Stan's poisson_lpmf(n | mu)  (functions.stan:19,33).

    n log(mu) - mu - lgamma(n + 1); mu == 0 and n == 0 contributes 0; mu == 0 and n > 0 gives -inf.
    (Stan raises for mu < 0 or NaN; here that propagates as NaN.)
    """
    n = np.asarray(n, dtype=np.float64)
    mu = np.asarray(mu, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        term = np.where(n > 0, n * np.log(mu), 0.0) - mu - gammaln(n + 1.0)
    return float(np.sum(term))


def loglik_bins(counts, time, exposure, omega1, omega2, beta1, beta2, dt, bkg, scale1, scale2, amplitude):
    """functions.stan:23-35 (``partial_log_like_bw_multi_scale_fast`` over the whole slice).

    poisson_lpmf(counts | exposure .* (exp((s1 cos tw1 + s2 cos tw2) b1 + (s1 sin tw1 + s2 sin tw2) b2) * A + B)),
    tw_i = (time - dt) (x) omega_i.
    """
    f = rff_multiscale(np.asarray(time, np.float64) - dt, omega1, omega2, beta1, beta2, 1.0, scale1, scale2)
    mu = np.asarray(exposure, np.float64) * (f * amplitude + bkg)
    return poisson_lpmf(counts, mu)


def total_loglik(counts, time, exposure, n_time_bins, omega1, omega2, beta1, beta2, dt, bkg, scale1, scale2, amplitude):
    """functions.stan:37-61.  Sum over detectors of loglik_bins on the un-padded prefix ``[:N_time_bins[d]]``.

    ``dt``, ``bkg``, ``amplitude`` are length-D (detector 0: dt = 0, amplitude = 1; rff_omega_ipn.stan:144-145).
    """
    lp = 0.0
    for d in range(len(n_time_bins)):
        nb = int(n_time_bins[d])
        lp += loglik_bins(
            counts[d][:nb], time[d][:nb], exposure[d][:nb], omega1, omega2, beta1, beta2,
            dt[d], bkg[d], scale1, scale2, amplitude[d],
        )
    return lp


def loglik_folded(counts, time, exposure, omega, a, b, dt, bkg, amplitude, with_lgamma=True):
    """Same as loglik_bins but on the folded (omega, a, b) form."""
    logf = log_intensity_folded(np.asarray(time, np.float64) - dt, omega, a, b)
    mu = np.asarray(exposure, np.float64) * (np.exp(logf) * amplitude + bkg)
    n = np.asarray(counts, np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        term = np.where(n > 0, n * np.log(mu), 0.0) - mu
    out = float(np.sum(term))
    if with_lgamma:
        out -= float(np.sum(gammaln(n + 1.0)))
    return out


def loglik_delay_grid(counts, time, exposure, omega, a, b, amp, bkg, delays):
    """The north-star grid: ll[g, s] for folded draws omega/a/b (S, J), amp/bkg (S,), delays (G,).

    Direct evaluation (no angle-addition shortcut), float64, one (delay, draw) at a time.
    """
    omega = np.atleast_2d(omega)
    a = np.atleast_2d(a)
    b = np.atleast_2d(b)
    S = omega.shape[0]
    G = len(delays)
    ll = np.empty((G, S))
    for s in range(S):
        for g in range(G):
            ll[g, s] = loglik_folded(counts, time, exposure, omega[s], a[s], b[s], delays[g], bkg[s], amp[s])
    return ll


def time_delay(grb_xyz, sc_pos_diff):
    """functions.stan:164-169.  dot(grb_xyz, sc_pos_diff) / c, positions in km -> seconds."""
    return np.dot(np.asarray(sc_pos_diff, np.float64), np.asarray(grb_xyz, np.float64)) * (1.0 / C_KM_S)


def sky_delays(ghat, sc_pos):
    """rff_omega_ipn.stan:63-68,129-134,144-145.  delays[p, d] = ghat_p . (p_0 - p_d) / c, d = 0 -> 0."""
    ghat = np.atleast_2d(np.asarray(ghat, np.float64))
    sc_pos = np.asarray(sc_pos, np.float64)
    diffs = sc_pos[0][None, :] - sc_pos  # (D, 3); row 0 is zero
    return (ghat @ diffs.T) * (1.0 / C_KM_S)


def loglik_sky_grid(counts, time, exposure, n_time_bins, sc_pos, ghat, omega, a, b, amp, bkg):
    """ll[p, s] = sum_d loglik(detector d shifted by delays[p, d]); amp/bkg are (D, S) (amp[0] must be 1)."""
    omega = np.atleast_2d(omega)
    a = np.atleast_2d(a)
    b = np.atleast_2d(b)
    S = omega.shape[0]
    dl = sky_delays(ghat, sc_pos)
    P, D = dl.shape
    ll = np.zeros((P, S))
    for s in range(S):
        for p in range(P):
            acc = 0.0
            for d in range(D):
                nb = int(n_time_bins[d])
                acc += loglik_folded(
                    counts[d][:nb], time[d][:nb], exposure[d][:nb], omega[s], a[s], b[s], dl[p, d], bkg[d][s], amp[d][s]
                )
            ll[p, s] = acc
    return ll


def grb_angles(grb_xyz):
    """rff_omega_ipn.stan:227-233.  phi = atan2(y, x); theta = -(acos(z) - pi/2)."""
    g = np.asarray(grb_xyz, np.float64)
    return np.arctan2(g[..., 1], g[..., 0]), -(np.arccos(g[..., 2]) - 0.5 * np.pi)


# ----------------------------------------------------------------------------------------------
# what a "posterior draw" is  (pyipn/stan_models/rff_omega_ipn.stan:100-124, 152-177)
# ----------------------------------------------------------------------------------------------
def draw_prior(rng, k, max_range, n_detectors=2):
    """One draw of the live model's parameters from its priors, transformed as the Stan model does.

    raw_scale positive_ordered[2] ~ N(1,1); scale = raw_scale / sqrt(k); range1_raw in (0,1) ~ LogN(0,.2);
    range2_raw in (0, range1_raw) ~ LogN(log .01, .2); bw = 1 / (range_raw * max_range); omega_i = omega_var_i * bw_i,
    omega_var ~ N(0,1); beta ~ N(0,1); log_bkg ~ N(log 500, log 100); log_amplitude ~ N(0,1).
    """
    while True:
        raw = np.sort(rng.normal(1.0, 1.0, size=2))
        if raw[0] > 0:
            break
    while True:
        r1 = rng.lognormal(0.0, 0.2)
        if 0 < r1 < 1:
            break
    while True:
        r2 = rng.lognormal(math.log(1e-2), 0.2)
        if 0 < r2 < r1:
            break
    scale = raw / math.sqrt(k)
    bw = 1.0 / (np.array([r1, r2]) * max_range)
    omega_var = rng.normal(size=(2, k))
    omega = omega_var * bw[:, None]
    beta1 = rng.normal(size=k)
    beta2 = rng.normal(size=k)
    log_bkg = rng.normal(math.log(500.0), math.log(100.0), size=n_detectors)
    log_amp = rng.normal(size=n_detectors - 1)
    return dict(
        beta1=beta1, beta2=beta2, omega=omega, omega_var=omega_var, bw=bw, scale=scale,
        bkg=np.exp(log_bkg), amplitude=np.exp(log_amp),
    )


# ----------------------------------------------------------------------------------------------
# data contract  (pyipn/lightcurve.py:24-47, pyipn/universe.py:441-529)
# ----------------------------------------------------------------------------------------------
def bin_events(arrival_times, tstart, tstop, dt):
    """lightcurve.py:24-47.  bins = arange(tstart, tstop, dt); histogram; rate = counts / dt."""
    bins = np.arange(tstart, tstop, dt)
    counts, edges = np.histogram(arrival_times, bins=bins)
    return counts / float(dt), edges, counts


def to_stan_data(arrival_times_per_det, sc_pos, tstart, tstop, dt=0.2, k=50, n_cores=1, factor=0.5, factor2=1):
    """universe.py:441-529 minus the astropy objects: padded (D, max_bins) counts / mid-times / exposures."""
    n_dets = len(arrival_times_per_det)
    tstart = np.atleast_1d(tstart)
    tstop = np.atleast_1d(tstop)
    dt = np.atleast_1d(dt)
    counts, times, exposures, n_time_bins = [], [], [], []
    for n in range(n_dets):
        m = n if n < len(tstart) else 0
        _, t, c = bin_events(arrival_times_per_det[n], tstart[m], tstop[m], dt[m])
        counts.append(c)
        times.append(np.mean([t[:-1], t[1:]], axis=0))
        exposures.append(t[1:] - t[:-1])
        n_time_bins.append(len(c))
    mx = max(n_time_bins)
    counts_stan = np.zeros((n_dets, mx), dtype=int)
    times_stan = np.zeros((n_dets, mx))
    exposure_stan = np.zeros((n_dets, mx))
    for n in range(n_dets):
        counts_stan[n, : n_time_bins[n]] = counts[n]
        times_stan[n, : n_time_bins[n]] = times[n]
        exposure_stan[n, : n_time_bins[n]] = exposures[n]
    grainsize = [int(np.round(n / n_cores) * factor) for n in n_time_bins]
    return dict(
        N_detectors=n_dets, N_time_bins=n_time_bins, max_N_time_bins=mx, counts=counts_stan, time=times_stan,
        exposure=exposure_stan, sc_pos=np.asarray(sc_pos, np.float64), k=k, grainsize=grainsize,
        grainsize_meta=max(1, int(np.round(n_dets / n_cores) * factor2)),
    )


# ----------------------------------------------------------------------------------------------
# HEALPix RING pix2vec (published algorithm, Gorski et al. 2005) -- sky grid only
# ----------------------------------------------------------------------------------------------
def healpix_ring_pix2vec(nside, ipix):
    """Unit vectors of RING-ordered HEALPix pixel centres; pure-Python loop, small/medium nside only."""
    ipix = np.atleast_1d(np.asarray(ipix, dtype=np.int64))
    npix = 12 * nside * nside
    ncap = 2 * nside * (nside - 1)
    out = np.empty((len(ipix), 3))
    for m, p in enumerate(ipix):
        p = int(p)
        if p < ncap:  # north polar cap
            iring = int((1 + math.isqrt(1 + 2 * p)) // 2)
            iphi = (p + 1) - 2 * iring * (iring - 1)
            z = 1.0 - iring * iring * 4.0 / npix
            phi = (iphi - 0.5) * math.pi / (2.0 * iring)
        elif p < npix - ncap:  # equatorial belt
            ip = p - ncap
            iring = ip // (4 * nside) + nside
            iphi = ip % (4 * nside) + 1
            fodd = 1.0 if ((iring + nside) & 1) else 0.5
            z = (2 * nside - iring) * 2.0 / (3.0 * nside)
            phi = (iphi - fodd) * math.pi / (2.0 * nside)
        else:  # south polar cap
            ip = npix - p
            iring = int((1 + math.isqrt(2 * ip - 1)) // 2)
            iphi = 4 * iring + 1 - (ip - 2 * iring * (iring - 1))
            z = -1.0 + iring * iring * 4.0 / npix
            phi = (iphi - 0.5) * math.pi / (2.0 * iring)
        st = math.sqrt(max(0.0, (1.0 - z) * (1.0 + z)))
        out[m] = (st * math.cos(phi), st * math.sin(phi), z)
    return out


# ----------------------------------------------------------------------------------------------
# classical IPN chi^2 cross-correlation  (pyipn/correlation.py:247-372)
# ----------------------------------------------------------------------------------------------
def correlate(arr_t_1, arr_cnts_1, arr_cnts_err_1, arr_t_2, arr_cnts_2, arr_cnts_err_2, i_beg_2, i_end_2, i_beg_1,
              n_max_1, fscale, dt_1, dt_2, n_sum_2=1):
    """correlation.py:247-372, plain-Python restatement (float64, same operation order): chi^2/dof of light curve 2
    against light curve 1 shifted by i = i_beg_1 .. i_beg_1 + n_max_1 - 1 bins with fractional re-binning.

    Returns (arr_dt, arr_chi, nDOF, fRijMin, dTmin, iMin, nMin) like the reference (first minimum wins)."""
    arr_dt = np.zeros(n_max_1)
    arr_chi = np.zeros(n_max_1)
    nDOF = (i_end_2 - i_beg_2) / n_sum_2
    fRijMin, iMin, nMin, dTmin = 1e4, 0, 0, 0.0
    n = 0
    for i in range(i_beg_1, i_beg_1 + n_max_1):
        fRij = 0.0
        fT1, fT2 = float(dt_1), 0.0
        l = i
        fdCnt1 = fdErr1 = 0.0
        for k in range(i_beg_2, i_end_2 + 1, n_sum_2):
            fCnt2 = fErr2 = 0.0
            for iSum2 in range(k, k + n_sum_2):
                fCnt2 += arr_cnts_2[iSum2]
                fErr2 += arr_cnts_err_2[iSum2]
            fCnt2 *= fscale
            fErr2 *= fscale * fscale
            fCnt1, fErr1 = fdCnt1, fdErr1
            while fT1 <= (fT2 + dt_2 * n_sum_2):
                fCnt1 += arr_cnts_1[l]
                fErr1 += arr_cnts_err_1[l]
                fT1 += dt_1
                l += 1
            fdT = (fT1 - (fT2 + dt_2 * n_sum_2)) / dt_1
            fdCnt1 = fdT * arr_cnts_1[l]
            fCnt1 += (1.0 - fdT) * arr_cnts_1[l]
            fdErr1 = fdT * arr_cnts_err_1[l]
            fErr1 += (1.0 - fdT) * arr_cnts_err_1[l]
            l += 1
            fT1 += dt_1
            fT2 += dt_2 * n_sum_2
            fDif = fCnt1 - fCnt2
            fRij += fDif * fDif / (fErr1 + fErr2)
        fRij /= nDOF - 1
        dT = arr_t_2[i_beg_2] - arr_t_1[i]
        arr_dt[n], arr_chi[n] = dT, fRij
        if fRij < fRijMin:
            fRijMin, iMin, nMin, dTmin = fRij, i, n, dT
        n += 1
    return arr_dt, arr_chi, nDOF, fRijMin, dTmin, iMin, nMin


def get_max_sn(counts, dt, dt_ms, bkg_rate=500.0):
    """lightcurve.py:146-173: window of len_int bins with the largest background-subtracted counts (first maximum)."""
    dt_ms_bin = dt * 1000.0
    len_int = int(dt_ms // dt_ms_bin)
    cs = np.cumsum(np.asarray(counts) - bkg_rate * dt)
    fmax, i1, i2 = 0.0, 0, cs.size
    for i in range(cs.size - len_int):
        dif = cs[i + len_int] - cs[i]
        if dif > fmax:
            fmax, i1, i2 = dif, i, i + len_int
    return i1, i2


def dtcc_interval(arr_dt, arr_chi, nDOF, fRijMin, nMin, nSigma):
    """correlation.py:127-171 (_get_dTcc): chi^2 confidence interval of the lag."""
    from scipy.stats import chi2, norm

    P0 = norm.cdf(nSigma) - norm.cdf(-nSigma)
    fSigma = chi2.ppf(P0, nDOF - 1) / (nDOF - 1) - 1.0 + fRijMin
    n = arr_chi.size
    i = nMin
    fRij = arr_chi[i]
    while (i > 1) and (fRij < fSigma):
        i -= 1
        fRij = arr_chi[i]
    dTupper = arr_dt[i]
    i = nMin
    fRij = arr_chi[i]
    while (i < n - 1) and (fRij < fSigma):
        i += 1
        fRij = arr_chi[i]
    return arr_dt[i], dTupper, fSigma


# ----------------------------------------------------------------------------------------------
# sky post-processing.  Parity UNPINNED: astropy / mocpy / ligo.skymap are not
# installed, so the reference functions cannot be run here; these restate their documented arithmetic.
# ----------------------------------------------------------------------------------------------
C_KM_S_ASTROPY = 299792.458  # astropy.constants.c in km/s, what utils/timing.py uses (the Stan models use 299792.46)


def theta_from_time_delay(dt, distance):
    """utils/timing.py:14-23.  Annulus half-opening angle: arccos(round(c dt / distance, 15)), dt in s, distance in km."""
    return np.arccos(np.around(C_KM_S_ASTROPY * np.asarray(dt, np.float64) / np.asarray(distance, np.float64), 15))


def calculate_distance_and_norm(p1, p2):
    """utils/timing.py:26-59 on plain km vectors: d = p2 - p1 -> (|d|, d/|d|, ra, dec) with ra in [0, 2 pi)
    (astropy's longitude wrap) and dec = asin(z/|d|), both in radians."""
    d = np.asarray(p2, np.float64) - np.asarray(p1, np.float64)
    distance = np.sqrt(np.sum(d * d))
    n = d / distance
    ra = np.arctan2(n[1], n[0]) % (2.0 * np.pi)
    dec = np.arcsin(n[2])
    return distance, n, ra, dec


def localize_grb(sc_pos, T0):
    """universe.py:670-697 (localize_GRB) with calculate_annulus (:364-386): for every detector pair (i < j) the row
    M = unit vector from i to j and b = cos(theta_ij), theta from dt = T0[i] - T0[j]; least squares M g = b."""
    from itertools import combinations

    sc_pos = np.asarray(sc_pos, np.float64)
    M, b = [], []
    for i, j in combinations(range(sc_pos.shape[0]), 2):
        distance, n, _, _ = calculate_distance_and_norm(sc_pos[i], sc_pos[j])
        theta = theta_from_time_delay(T0[i] - T0[j], distance)
        M.append(n)
        b.append(np.cos(theta))
    g = np.linalg.lstsq(np.array(M), np.array(b), rcond=None)[0]
    return g


def sky_credible(ll, log_weight=None):
    """Greedy credible levels on a pixel grid, the grid-native replacement of fit.py:193-206 + :328-347
    (KDE -> multi-order map -> MOC.from_valued_healpix_cells(..., cumul_to=c)).  ll (P,) or (P, S)."""
    ll = np.asarray(ll, np.float64)
    if ll.ndim == 1:
        ll = ll[:, None]
    P, S = ll.shape
    m = ll.max(axis=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        L = np.where(np.isfinite(m), m + np.log(np.sum(np.exp(ll - np.where(np.isfinite(m), m, 0.0)[:, None]), axis=1)) - np.log(S), m)
    if log_weight is not None:
        L = L + np.asarray(log_weight, np.float64)
    mx = L.max()
    e = np.exp(L - mx)
    tot = e.sum()
    prob = e / tot
    order = np.argsort(-prob, kind="stable")
    level = np.empty(P)
    level[order] = np.cumsum(prob[order])
    return prob, level, order, mx + np.log(tot)
