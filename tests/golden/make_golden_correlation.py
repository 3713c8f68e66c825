"""Mint `correlation_golden.npz` from the REFERENCE's own numba `correlate` and `Correlator`
(pyipn/correlation.py:9-372, pyipn/lightcurve.py:96-180), build container only.

`import pyipn` needs astropy/matplotlib/arviz; the two modules used here only touch `matplotlib.pyplot` (never called)
and `astropy.units` (one seconds->milliseconds conversion), so both are stubbed and the reference files are executed
unmodified through a synthetic `pyipn` package whose __path__ is the reference directory.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/pyipn"


def import_reference():
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nbcache")
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)

    class _Q:
        def __init__(self, v, scale=1.0):
            self.v, self.scale = v, scale

        def __rmul__(self, other):
            return _Q(other, self.scale)

        def to(self, unit):
            assert unit == "millisecond"
            return _Q(self.v * self.scale * 1000.0, 1.0)

        @property
        def value(self):
            return self.v

    astropy = types.ModuleType("astropy")
    units = types.ModuleType("astropy.units")
    units.s = _Q(1.0, 1.0)
    astropy.units = units
    sys.modules.setdefault("astropy", astropy)
    sys.modules.setdefault("astropy.units", units)
    pkg = types.ModuleType("pyipn")
    pkg.__path__ = [REF]
    sys.modules["pyipn"] = pkg
    import importlib

    lc = importlib.import_module("pyipn.lightcurve")
    corr = importlib.import_module("pyipn.correlation")
    return lc, corr


def norris(x, K, t_start, t_rise, t_decay):
    out = np.zeros_like(x)
    m = x > t_start
    d = x[m] - t_start
    out[m] = K * np.exp(2.0 * np.sqrt(t_rise / t_decay)) * np.exp((-t_rise / d) - d / t_decay)
    return out


def main():
    lc_mod, corr_mod = import_reference()
    rng = np.random.default_rng(20261020)
    out = {}
    # two detectors, binned at different resolutions (lc1 finer), true lag 0.23 s
    tstart, tstop = -10.0, 30.0
    for name, dt, lag, seed in (("1", 0.064, 0.0, 1), ("2", 0.256, 0.23, 2)):
        edges = np.arange(tstart, tstop, dt)
        mid = 0.5 * (edges[:-1] + edges[1:])
        rate = norris(mid - lag, 4000.0, 0.0, 0.5, 4.0) + 500.0
        counts = np.random.default_rng(seed).poisson(rate * dt)
        out["edges" + name], out["counts" + name], out["dt" + name] = edges, counts.astype(np.int64), np.float64(dt)
    b1 = lc_mod.BinnedLightCurve(out["counts1"], out["edges1"], tstart, tstop, float(out["dt1"]))
    b2 = lc_mod.BinnedLightCurve(out["counts2"], out["edges2"], tstart, tstop, float(out["dt2"]))
    i_beg2, i_end2 = int(b2.time2idx(-1.0)), int(b2.time2idx(12.0))
    idx1 = int(b1.time2idx(-6.0))
    c = corr_mod.Correlator(b1, b2, idx1, i_beg2, i_end2, cl_sigma=[1, 2, 3], bkg_rate=500)
    out.update(idx_lc_1=idx1, idx_beg_lc2=i_beg2, idx_end_lc2=i_end2, n_max=c._n_max, fscale=c._fscale,
               idx_beg_lc1=c._idx_beg_lc1, idx_end_lc1=c._idx_end_lc1, arr_dt=c._arr_dt, arr_chi=c._arr_chi, nDOF=c._nDOF,
               fRijMin=c._fRijMin, dTmin=c._dTmin, iMin=c._iMin, nMin=c._nMin, dTlower=np.array(c._dTlower),
               dTupper=np.array(c._dTupper), fSigma=np.array(c._fSigma), res_ms1=b1.res_ms, res_ms2=b2.res_ms)
    # direct call with n_sum_2 = 2 and equal resolutions (integer re-binning path)
    r = corr_mod.correlate(b1.time_bins, b1.get_src_counts(500.0), b1.counts.astype(np.float64), b1.time_bins,
                           b1.get_src_counts(500.0) * 0.9, b1.counts.astype(np.float64), 150, 349, 20, 60, 1.1, b1.res_ms,
                           b1.res_ms, 2)
    out.update(B_arr_dt=r[0], B_arr_chi=r[1], B_nDOF=r[2], B_fRijMin=r[3], B_dTmin=r[4], B_iMin=r[5], B_nMin=r[6])
    np.savez_compressed(os.path.join(HERE, "correlation_golden.npz"), **out)
    print("dTmin", c._dTmin, "nMin", c._nMin, "n_max", c._n_max, "bounds", c._dTlower, c._dTupper)


if __name__ == "__main__":
    main()
