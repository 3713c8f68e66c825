"""Drop-in for the numeric part of ``pyipn/correlation.py`` (the classical IPN chi^2 cross-correlation of
Pal'shin et al. 2013): ``correlate`` keeps the reference's signature and return tuple,
``Correlator`` its constructor arguments and attributes; plotting is out of scope.
"""
from __future__ import annotations

import numpy as np

from ._lib import default_context
from .lightcurve import BinnedLightCurve


def correlate(arr_t_1, arr_cnts_1, arr_cnts_err_1, arr_t_2, arr_cnts_2, arr_cnts_err_2, i_beg_2, i_end_2, i_beg_1, n_max_1,
              fscale, dt_1, dt_2, n_sum_2=1):
    """pyipn/correlation.py:247-372 on the GPU (one thread per trial shift).
    Returns (arr_dt, arr_chi, nDOF, fRijMin, dTmin, iMin, nMin); the first minimum wins, like the reference's scan."""
    arr_dt, arr_chi = default_context().correlate(arr_t_1, arr_cnts_1, arr_cnts_err_1, arr_t_2, arr_cnts_2, arr_cnts_err_2,
                                                  i_beg_2, i_end_2, i_beg_1, n_max_1, fscale, dt_1, dt_2, n_sum_2)
    nDOF = (i_end_2 - i_beg_2) / n_sum_2
    fRijMin, iMin, nMin, dTmin = 1e4, 0, 0, 0.0  # correlation.py:296-299: only values below 1e4 register as a minimum
    if arr_chi.size:
        nbest = int(np.argmin(arr_chi))  # argmin returns the first occurrence, as `if fRij < fRijMin` does
        if arr_chi[nbest] < fRijMin:
            fRijMin, nMin, iMin, dTmin = float(arr_chi[nbest]), nbest, int(i_beg_1 + nbest), float(arr_dt[nbest])
    return arr_dt, arr_chi, nDOF, fRijMin, dTmin, iMin, nMin


class Correlator(object):
    """pyipn/correlation.py:9-171 without the plotting methods."""

    def __init__(self, lc_1, lc_2, idx_lc_1, idx_beg_lc2, idx_end_lc2, cl_sigma=[1, 2, 3], bkg_rate=500):
        assert isinstance(lc_1, BinnedLightCurve)
        assert isinstance(lc_2, BinnedLightCurve)
        self._idx_beg_lc2, self._idx_end_lc2 = idx_beg_lc2, idx_end_lc2
        self.lc_1, self.lc_2 = lc_1, lc_2
        self._idx_lc_1 = idx_lc_1
        dt_grb_ms = (idx_end_lc2 - idx_beg_lc2) * self.lc_2.res_ms
        self._idx_beg_lc1, self._idx_end_lc1 = self.lc_1.get_max_sn(dt_grb_ms, bkg_rate)
        self._fscale = np.sum(lc_1.get_src_counts()[self._idx_beg_lc1: self._idx_end_lc1 + 1]) / np.sum(
            lc_2.get_src_counts()[self._idx_beg_lc2: self._idx_end_lc2 + 1])
        self._n_max = int(lc_1.n_bins - idx_lc_1 - (idx_end_lc2 - idx_beg_lc2 + 1) * lc_2.res_ms // lc_1.res_ms)
        self._cl_sigma = cl_sigma
        (self._arr_dt, self._arr_chi, self._nDOF, self._fRijMin, self._dTmin, self._iMin, self._nMin) = correlate(
            lc_1.time_bins, lc_1.get_src_counts(bkg_rate), np.asarray(lc_1.counts, dtype=np.float64), lc_2.time_bins,
            lc_2.get_src_counts(bkg_rate), np.asarray(lc_2.counts, dtype=np.float64), idx_beg_lc2, idx_end_lc2, idx_lc_1,
            self._n_max, self._fscale, lc_1.res_ms, lc_2.res_ms)
        self._dTlower, self._dTupper, self._fSigma = [], [], []
        for sigma in self._cl_sigma:
            lo, up, fs = self._get_dTcc(nSigma=sigma)
            self._dTlower.append(lo)
            self._dTupper.append(up)
            self._fSigma.append(fs)

    def _get_dTcc(self, nSigma=3):
        """Lag confidence interval at ``nSigma`` (semantics of correlation.py:127-171): the chi^2/dof level that the two-sided
        Gaussian probability of nSigma maps to, and the first shifts on either side of the minimum at which the curve reaches it
        -- found with two vectorised threshold searches instead of the reference's two walks.  The reference's walks stop at
        index 1 on the left and at the last index on the right even when the curve never reaches the level there: kept."""
        from scipy.stats import chi2, norm

        dof = self._nDOF - 1
        level = chi2.isf(2.0 * norm.sf(nSigma), dof) / dof - 1.0 + self._fRijMin
        above = np.asarray(self._arr_chi) >= level
        k = self._nMin
        left = np.flatnonzero(above[2:k + 1]) + 2 if k >= 2 else np.empty(0, dtype=np.int64)
        i_upper = int(left[-1]) if left.size else min(k, 1)
        right = np.flatnonzero(above[k:-1]) + k
        i_lower = int(right[0]) if right.size else max(k, above.size - 1)
        return self._arr_dt[i_lower], self._arr_dt[i_upper], level

    def info(self):
        """correlation.py:98-125 (the summary the reference prints); returned as well as printed."""
        dT2 = (self._idx_end_lc2 - self._idx_beg_lc2) * self.lc_2.res_ms
        dT1 = (self._idx_end_lc1 - self._idx_beg_lc1) * self.lc_1.res_ms
        out = "Cross-correlation info:\n"
        out += "CC interval for lc2 (i_beg, i_end, dT): {:4d} {:4d} {:4d} ms\n".format(self._idx_beg_lc2, self._idx_end_lc2, int(dT2))
        out += "CC interval for lc1 (i_beg, i_end, dT): {:4d} {:4d} {:4d} ms\n".format(self._idx_beg_lc1, self._idx_end_lc1, int(dT1))
        out += "Lc scale factor: {:.3f}\n".format(self._fscale)
        out += "Lc 1 i_start: {:d}\n".format(self._idx_lc_1)
        out += "Lc 1 n_max: {:d}\n".format(self._n_max)
        for i, sigma in enumerate(self._cl_sigma):
            out += "Calculated time delay and its {:.1f} sigma CI:\n".format(sigma)
            out += "dTcc dTcc- dTcc+: {:.3f} {:+.3f} {:+.3f}\n".format(self._dTmin, self._dTlower[i] - self._dTmin,
                                                                      self._dTupper[i] - self._dTmin)
        print(out)
        return out

    # correlation.py:173-187
    dt_min = property(lambda self: self._dTmin)
    dt_lower = property(lambda self: self._dTlower)
    dt_upper = property(lambda self: self._dTupper)
    f_sigma = property(lambda self: self._fSigma)
