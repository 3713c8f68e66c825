"""Next-row N1: the classical chi^2 cross-correlation (pyipn/correlation.py:247-372).

CPU: the oracle restatement against the fixture minted from the reference's own numba `correlate` / `Correlator`
(tests/golden/make_golden_correlation.py).  GPU: the CUDA kernel and the `Correlator` facade against the same fixture."""
import os

import numpy as np
import pytest

from oracle import ipn_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "correlation_golden.npz")


@pytest.fixture(scope="module")
def g():
    return dict(np.load(GOLDEN))


def _args(g, bkg=500.0):
    c1, c2 = g["counts1"].astype(np.float64), g["counts2"].astype(np.float64)
    return (g["edges1"], c1 - bkg * float(g["dt1"]), c1, g["edges2"], c2 - bkg * float(g["dt2"]), c2, int(g["idx_beg_lc2"]),
            int(g["idx_end_lc2"]), int(g["idx_lc_1"]), int(g["n_max"]), float(g["fscale"]), float(g["res_ms1"]), float(g["res_ms2"]))


def test_oracle_correlate_matches_reference(g):
    arr_dt, arr_chi, nDOF, fmin, dTmin, iMin, nMin = orc.correlate(*_args(g))
    np.testing.assert_allclose(arr_dt, g["arr_dt"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(arr_chi, g["arr_chi"], rtol=1e-11)
    assert (nDOF, iMin, nMin) == (float(g["nDOF"]), int(g["iMin"]), int(g["nMin"]))
    assert fmin == pytest.approx(float(g["fRijMin"]), rel=1e-11) and dTmin == pytest.approx(float(g["dTmin"]), abs=1e-12)
    # n_sum_2 = 2, equal resolutions
    c1 = g["counts1"].astype(np.float64)
    src = c1 - 500.0 * float(g["dt1"])
    r = orc.correlate(g["edges1"], src, c1, g["edges1"], src * 0.9, c1, 150, 349, 20, 60, 1.1, float(g["res_ms1"]),
                      float(g["res_ms1"]), 2)
    np.testing.assert_allclose(r[1], g["B_arr_chi"], rtol=1e-11)
    assert r[6] == int(g["B_nMin"]) and r[2] == float(g["B_nDOF"])


def test_oracle_max_sn_and_interval_match_reference(g):
    dt_grb_ms = (int(g["idx_end_lc2"]) - int(g["idx_beg_lc2"])) * float(g["res_ms2"])
    assert orc.get_max_sn(g["counts1"], float(g["dt1"]), dt_grb_ms, 500) == (int(g["idx_beg_lc1"]), int(g["idx_end_lc1"]))
    for i, s in enumerate((1, 2, 3)):
        lo, up, fs = orc.dtcc_interval(g["arr_dt"], g["arr_chi"], float(g["nDOF"]), float(g["fRijMin"]), int(g["nMin"]), s)
        assert lo == pytest.approx(g["dTlower"][i]) and up == pytest.approx(g["dTupper"][i]) and fs == pytest.approx(g["fSigma"][i])


def test_correlator_host_logic_without_gpu(g, monkeypatch, capsys):
    """Everything in `Correlator` around the chi^2 scan (max-S/N window, scale factor, n_max, confidence intervals, the
    reference's accessors and `info`, correlation.py:16-187) against the fixture, with the device call replaced by the
    oracle's scan (host logic only; the kernel itself is checked in the GPU tests below)."""
    import pyipn_b200.correlation as cm
    from pyipn_b200 import BinnedLightCurve, Correlator

    class Fake:
        def correlate(self, *args):
            r = orc.correlate(*args)
            return r[0], r[1]

    monkeypatch.setattr(cm, "default_context", lambda: Fake())
    b1 = BinnedLightCurve(g["counts1"], g["edges1"], -10.0, 30.0, float(g["dt1"]))
    b2 = BinnedLightCurve(g["counts2"], g["edges2"], -10.0, 30.0, float(g["dt2"]))
    c = Correlator(b1, b2, int(g["idx_lc_1"]), int(g["idx_beg_lc2"]), int(g["idx_end_lc2"]), cl_sigma=[1, 2, 3], bkg_rate=500)
    assert (c._idx_beg_lc1, c._idx_end_lc1, c._n_max) == (int(g["idx_beg_lc1"]), int(g["idx_end_lc1"]), int(g["n_max"]))
    assert c.dt_min == pytest.approx(float(g["dTmin"]), abs=1e-12)
    assert len(c.dt_lower) == len(c.dt_upper) == 3  # the reference test's own assertion (test_basic.py:92-93)
    np.testing.assert_allclose(c.dt_lower, g["dTlower"], atol=1e-12)
    np.testing.assert_allclose(c.dt_upper, g["dTupper"], atol=1e-12)
    np.testing.assert_allclose(c.f_sigma, g["fSigma"], rtol=1e-10)
    text = c.info()
    assert "Cross-correlation info" in text and text.count("sigma CI") == 3 and "Lc 1 n_max: %d" % int(g["n_max"]) in text
    assert capsys.readouterr().out.strip() == text.strip()


@pytest.mark.gpu
def test_gpu_correlate_matches_reference(ctx, g):
    from pyipn_b200 import correlate

    arr_dt, arr_chi, nDOF, fmin, dTmin, iMin, nMin = correlate(*_args(g))
    np.testing.assert_allclose(arr_dt, g["arr_dt"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(arr_chi, g["arr_chi"], rtol=1e-11)
    assert (iMin, nMin) == (int(g["iMin"]), int(g["nMin"]))  # arg-min exact
    assert dTmin == pytest.approx(float(g["dTmin"]), abs=1e-12) and nDOF == float(g["nDOF"])
    c1 = g["counts1"].astype(np.float64)
    src = c1 - 500.0 * float(g["dt1"])
    r = correlate(g["edges1"], src, c1, g["edges1"], src * 0.9, c1, 150, 349, 20, 60, 1.1, float(g["res_ms1"]),
                  float(g["res_ms1"]), 2)
    np.testing.assert_allclose(r[1], g["B_arr_chi"], rtol=1e-11)
    assert r[6] == int(g["B_nMin"])


@pytest.mark.gpu
def test_gpu_correlator_facade_matches_reference(ctx, g):
    from pyipn_b200 import BinnedLightCurve, Correlator

    b1 = BinnedLightCurve(g["counts1"], g["edges1"], -10.0, 30.0, float(g["dt1"]))
    b2 = BinnedLightCurve(g["counts2"], g["edges2"], -10.0, 30.0, float(g["dt2"]))
    c = Correlator(b1, b2, int(g["idx_lc_1"]), int(g["idx_beg_lc2"]), int(g["idx_end_lc2"]), cl_sigma=[1, 2, 3], bkg_rate=500)
    assert (c._idx_beg_lc1, c._idx_end_lc1, c._n_max) == (int(g["idx_beg_lc1"]), int(g["idx_end_lc1"]), int(g["n_max"]))
    assert c._fscale == pytest.approx(float(g["fscale"]), rel=1e-13)
    assert c.dt_min == pytest.approx(float(g["dTmin"]), abs=1e-12)
    assert len(c._dTlower) == len(c._dTupper) == 3  # the reference test's own assertion (test_basic.py:92-93)
    np.testing.assert_allclose(c._dTlower, g["dTlower"], atol=1e-12)
    np.testing.assert_allclose(c._dTupper, g["dTupper"], atol=1e-12)
    np.testing.assert_allclose(c._fSigma, g["fSigma"], rtol=1e-10)


@pytest.mark.gpu
def test_gpu_correlate_large_random(ctx):
    """20000 shifts x 2000 bins, ragged resolutions: GPU vs oracle on a subsample of shifts, arg-min exact."""
    rng = np.random.default_rng(4)
    n1, n2 = 60_000, 2_600
    dt1, dt2 = 0.002, 0.0315
    t1, t2 = np.arange(n1 + 1) * dt1, np.arange(n2 + 1) * dt2
    prof = lambda t: 300.0 * np.exp(-0.5 * ((t - 40.0) / 6.0) ** 2) * (1 + 0.4 * np.sin(t * 1.7))
    c1 = rng.poisson((prof(t1[:-1] + 3.3) + 500.0) * dt1).astype(np.float64)
    c2 = rng.poisson((prof(t2[:-1]) + 500.0) * dt2).astype(np.float64)
    s1, s2 = c1 - 500.0 * dt1, c2 - 500.0 * dt2
    args = (t1, s1, c1, t2, s2, c2, 300, 2299, 100, 20_000, 1.0, dt1 * 1000.0, dt2 * 1000.0)  # re-binned lc1 ~ lc2: scale 1
    arr_dt, arr_chi = ctx.correlate(*args)
    nbest = int(np.argmin(arr_chi))
    assert abs(arr_dt[nbest] - 3.3) < 0.1
    lo = max(0, nbest - 3)
    sub = orc.correlate(t1, s1, c1, t2, s2, c2, 300, 2299, 100 + lo, 7, 1.0, dt1 * 1000.0, dt2 * 1000.0)
    np.testing.assert_allclose(arr_chi[lo:lo + 7], sub[1], rtol=1e-11)
    assert lo + sub[6] == nbest
