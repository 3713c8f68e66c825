"""Host-side mirror of the hot-path part of ``pyipn/fit.py``.

* ``_expeced_rate`` / ``_expeced_rate_multiscale`` / ``ppc_generator`` keep the reference's (mis-spelt) names,
  argument order and array layouts (pyipn/fit.py:651-705) and run on the B200 through libipn_b200.
* ``Fit`` keeps the posterior-array contract of ``Fit.__init__`` (pyipn/fit.py:62-172) and
  ``expected_rate`` (pyipn/fit.py:237-305), and adds the north-star grid evaluations
  ``delay_grid_loglik`` / ``sky_grid_loglik``.  Plotting, KDE/MOC sky maps and the arviz loaders are out of
  scope; ``from_netcdf`` only works if arviz is importable.
"""
from __future__ import annotations

import numpy as np

from ._lib import default_context
from .params import fold_draws


def _expeced_rate(time, omega1, omega2, beta1, beta2, bw, scale, amplitude, dt, N):
    """pyipn/fit.py:651-668.  ``out[n] = amplitude[n] * RFF(time - dt[n], w1[:, n], ..., bw[n], scale[n])`` -> (N, T)."""
    omega, a, b = fold_draws(np.asarray(omega1)[:, :N], np.asarray(omega2)[:, :N], np.asarray(beta1)[:, :N],
                             np.asarray(beta2)[:, :N], np.asarray(bw)[:N], np.asarray(scale)[:N])
    return default_context().rff_eval(time, omega, a, b, np.asarray(dt, dtype=np.float64)[:N],
                                      np.asarray(amplitude, dtype=np.float64)[:N])


def _expeced_rate_multiscale(time, omega1, omega2, beta1, beta2, bw, scale, amplitude, dt, N):
    """pyipn/fit.py:671-691.  ``scale`` is (2, S): scale[0, n] weights the omega1 set, scale[1, n] the omega2 set."""
    omega, a, b = fold_draws(np.asarray(omega1)[:, :N], np.asarray(omega2)[:, :N], np.asarray(beta1)[:, :N],
                             np.asarray(beta2)[:, :N], np.asarray(bw)[:N], np.asarray(scale)[:, :N])
    return default_context().rff_eval(time, omega, a, b, np.asarray(dt, dtype=np.float64)[:N],
                                      np.asarray(amplitude, dtype=np.float64)[:N])


def ppc_generator(time, exposure, rates, bkg, N, seed=None):
    """pyipn/fit.py:694-705.  ``out[n, i] ~ Poisson((rates[n, i] + bkg[n]) * exposure[i])`` as float64 (N, T).

    The reference draws from numba's global MT19937 state; here each (n, i) has its own Philox stream keyed by
    ``seed`` (fresh entropy when None), so parity with the reference is statistical, not stream-exact."""
    if seed is None:
        seed = int(np.random.SeedSequence().generate_state(1, dtype=np.uint64)[0])
    rates = np.asarray(rates, dtype=np.float64)[:N]
    out = default_context().poisson_counts(exposure, rates, np.asarray(bkg, dtype=np.float64)[:N], seed)
    return out.astype(np.float64)


class Fit(object):
    """Posterior container with the reference's array layout (draw = fast axis).

    Parameters mirror the variables ``Fit.__init__`` pulls out of the arviz posterior (pyipn/fit.py:62-172):
    ``beta1, beta2 (k, S)``; ``omega (2, k, S)``; ``scale (S,)`` or ``(2, S)``; ``background (D, S)`` (or ``(S,)``);
    ``amplitude (D-1, S)``; ``dt (D-1, S)``; ``grb_theta, grb_phi (S,)``.
    """

    def __init__(self, beta1, beta2, omega, scale, background, amplitude=None, dt=None, grb_theta=None,
                 grb_phi=None, universe=None):
        self._beta1 = np.asarray(beta1, dtype=np.float64)
        self._beta2 = np.asarray(beta2, dtype=np.float64)
        self._n_samples = self._beta1.shape[-1]
        omega = np.asarray(omega, dtype=np.float64)
        self._omega1 = omega[0]
        self._omega2 = omega[1]
        self._scale = np.asarray(scale, dtype=np.float64)
        self._multi_scale = self._scale.ndim == 2 and self._scale.shape[0] == 2  # fit.py:100-106
        self._background = np.asarray(background, dtype=np.float64)
        self._amplitude = np.ones(self._n_samples) if amplitude is None else np.asarray(amplitude, dtype=np.float64)
        if dt is not None:  # fit.py:108-136
            self._dt = np.atleast_2d(np.asarray(dt, dtype=np.float64))
            self._amplitude = np.atleast_2d(self._amplitude)
            self._grb_theta = None if grb_theta is None else np.asarray(grb_theta, dtype=np.float64)
            self._grb_phi = None if grb_phi is None else np.asarray(grb_phi, dtype=np.float64)
            self._is_dt_fit = True
            self._n_dets = self._background.shape[0]
        else:
            self._dt = None
            self._grb_theta = None
            self._grb_phi = None
            self._is_dt_fit = False
            self._n_dets = 1
        self._use_bw = False  # fit.py:135: Fit always evaluates with bw = 1
        self._universe = universe
        self._has_universe = universe is not None

    # ---- constructors --------------------------------------------------------------------------------
    @classmethod
    def from_arrays(cls, **posterior):
        return cls(**posterior)

    @classmethod
    def from_netcdf(cls, *file_name, fast_open=False):
        """pyipn/fit.py:219-230; needs arviz (absent in the build image -> ImportError)."""
        import arviz as av  # noqa: F401  (deliberately not a hard dependency)

        idata = [av.from_netcdf(f) for f in file_name]
        post = idata[0] if len(idata) == 1 else av.concat(*idata, dim="chain")
        p = post.posterior

        def get(name):
            return getattr(p, name).stack(sample=("chain", "draw")).values

        kw = dict(beta1=get("beta1"), beta2=get("beta2"), omega=get("omega"), scale=get("scale"), background=get("bkg"))
        for name, key in (("amplitude", "amplitude"), ("dt", "dt"), ("grb_theta", "grb_theta"), ("grb_phi", "grb_phi")):
            try:
                kw[key] = get(name)
            except AttributeError:
                pass
        return cls(**kw)

    def set_universe(self, universe):
        self._universe = universe
        self._has_universe = True

    @property
    def n_samples(self):
        return self._n_samples

    # pyipn/fit.py:581-648: the posterior arrays in the reference's layout (draw = fast axis)
    beta1 = property(lambda self: self._beta1)
    beta2 = property(lambda self: self._beta2)
    omega1 = property(lambda self: self._omega1)
    omega2 = property(lambda self: self._omega2)
    scale = property(lambda self: self._scale)
    dt = property(lambda self: self._dt)
    grb_theta = property(lambda self: self._grb_theta)
    grb_phi = property(lambda self: self._grb_phi)
    amplitude = property(lambda self: self._amplitude)
    background = property(lambda self: self._background)

    def _detector_check(self, det_number):
        assert det_number in range(self._n_dets)

    # ---- pyipn/fit.py:237-305 ---------------------------------------------------------------------------
    def _dt_amp(self, detector):
        if self._is_dt_fit and detector > 0:
            dt = self._dt[detector - 1]
        else:
            dt = np.zeros(self._n_samples)
        if not self._is_dt_fit:
            amp = self._amplitude
        elif detector == 0:
            amp = np.ones(self._n_samples)
        else:
            amp = self._amplitude[detector - 1]
        return dt, amp

    def expected_rate(self, time, detector):
        self._detector_check(detector)
        dt, amp = self._dt_amp(detector)
        bw = np.ones(self._n_samples)
        fn = _expeced_rate_multiscale if self._multi_scale else _expeced_rate
        return fn(time, self._omega1, self._omega2, self._beta1, self._beta2, bw, self._scale, amp, dt, self._n_samples)

    def _background_of(self, detector):
        return self._background[detector] if self._is_dt_fit else self._background

    # ---- pyipn/fit.py:555-579 with explicit bins instead of the Universe's light curve ---------------------
    def compute_ppcs(self, detector, mid_points, exposure, seed=None):
        pred_rate = self.expected_rate(mid_points, detector)
        return ppc_generator(mid_points, exposure, pred_rate, self._background_of(detector), self._n_samples, seed=seed)

    # ---- the numeric halves of the reference's callers of the path (fit.py:447-579); plotting stays out of scope -------
    def _binned(self, detector, tstart, tstop, dt):
        self._detector_check(detector)
        assert self._has_universe
        lc = self._universe.light_curves[list(self._universe.light_curves.keys())[detector]]
        rate, edges, counts = lc.get_binned_light_curve(tstart, tstop, dt)
        return rate, edges, counts, 0.5 * (edges[:-1] + edges[1:]), np.diff(edges)

    def _compute_ppcs(self, detector, tstart, tstop, dt, seed=None):
        """fit.py:555-579: posterior-predictive counts (S, T) on the universe's light curve of ``detector``."""
        _, _, _, mid_points, exposure = self._binned(detector, tstart, tstop, dt)
        return self.compute_ppcs(detector, mid_points, exposure, seed=seed)

    def light_curve_fit(self, detector, tstart, tstop, dt=0.2, thin=1):
        """What plot_light_curve_fit draws (fit.py:447-484): ``(mid_points, observed rate, model rate + background)`` with
        the model as ``(S // thin, T)``."""
        rate, _, _, mid_points, _ = self._binned(detector, tstart, tstop, dt)
        pred = self.expected_rate(mid_points, detector)
        bkg = np.asarray(self._background_of(detector), dtype=np.float64)
        return mid_points, rate, (pred + bkg[:, None])[::thin]

    def light_curve_ppc_bands(self, detector, tstart, tstop, dt=0.2, levels=(99, 95, 68), seed=None):
        """What plot_light_curve_ppcs draws (fit.py:486-553): per level the (low, high) percentile bands of the
        posterior-predictive rate, plus the bin edges and the observed rate."""
        rate, edges, _, _, exposure = self._binned(detector, tstart, tstop, dt)
        ppcs = self._compute_ppcs(detector, tstart, tstop, dt, seed=seed) / exposure
        bands = [(np.percentile(ppcs, 50.0 - lv / 2.0, axis=0), np.percentile(ppcs, 50.0 + lv / 2.0, axis=0)) for lv in levels]
        return edges, rate, bands

    # ---- folded draws + north-star grids -------------------------------------------------------------------
    def folded(self):
        return fold_draws(self._omega1, self._omega2, self._beta1, self._beta2, np.ones(self._n_samples), self._scale)

    def delay_grid_loglik(self, detector, time, exposure, counts, delays):
        """ll[g, s] of detector ``detector``'s binned counts over trial delays x posterior draws."""
        from .likelihood import loglik_delay_grid

        self._detector_check(detector)
        _, amp = self._dt_amp(detector)
        omega, a, b = self.folded()
        return loglik_delay_grid(time, exposure, counts, omega, a, b, amp, self._background_of(detector), delays)

    def sky_grid_loglik(self, stan_data, ghat):
        """ll[p, s] summed over detectors for sky directions ``ghat`` (P, 3), data as Universe.to_stan_data gives."""
        from .likelihood import loglik_sky_grid

        D = stan_data["N_detectors"]
        amp = np.ones((D, self._n_samples))
        if self._is_dt_fit and D > 1:
            amp[1:] = self._amplitude[: D - 1]
        omega, a, b = self.folded()
        return loglik_sky_grid(stan_data["N_time_bins"], stan_data["time"], stan_data["exposure"], stan_data["counts"],
                               stan_data["sc_pos"], ghat, omega, a, b, amp, np.atleast_2d(self._background))

    def sky_credible(self, stan_data, ghat, levels=(0.68, 0.95), pixel_area=None):
        """Draw-marginal credible regions on the sky grid ``ghat`` (P, 3): ``sky_grid_loglik`` followed by
        ``pyipn_b200.sky.credible_regions`` (grid-native replacement of _build_moc_map / _contour_more_detectors,
        pyipn/fit.py:193-206, 328-347)."""
        from .sky import credible_regions

        return credible_regions(self.sky_grid_loglik(stan_data, ghat), levels=levels, pixel_area=pixel_area)
