"""Mint the golden fixtures in this directory from the REFERENCE's own code.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

What is pinned
--------------
``rff_golden.npz``      : outputs of the reference's numba ``RFF`` / ``RFF_multiscale`` (pyipn/rff.py:5-29), loaded by
                          file path (``import pyipn`` needs astropy, absent here), on seeded inputs, plus the draw loop of
                          pyipn/fit.py:651-691 executed around those reference functions (fit.py itself cannot import:
                          arviz/matplotlib absent, so the 6-line loop is re-run here calling the reference RFF).
``loglik_golden.npz``   : the composition  reference-RFF_multiscale -> scipy.stats.poisson.logpmf  on a small delay grid.
                          scipy stands in for Stan math's poisson_lpmf (un-vendored by the reference, parity unpinned).

The fixtures are small (< 200 KB) and are committed together with this script.
"""
import importlib.util
import os
import sys

import numpy as np
from scipy.stats import poisson

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/pyipn/rff.py"


def load_ref():
    spec = importlib.util.spec_from_file_location("ref_rff", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)  # rff.py has no cache=True -> writes nothing into /root/reference
    return mod


def make_params(rng, k, S, max_range):
    """(k, S) arrays in the layout Fit.__init__ produces (pyipn/fit.py:62-106)."""
    scale = np.sort(np.abs(rng.normal(1.0, 0.3, size=(2, S))), axis=0) / np.sqrt(k)
    r1 = np.clip(rng.lognormal(0.0, 0.2, size=S), 0.05, 0.999)
    r2 = np.minimum(rng.lognormal(np.log(1e-2), 0.2, size=S), 0.9 * r1)
    omega1 = rng.normal(size=(k, S)) / (r1 * max_range)
    omega2 = rng.normal(size=(k, S)) / (r2 * max_range)
    beta1 = rng.normal(size=(k, S))
    beta2 = rng.normal(size=(k, S))
    return omega1, omega2, beta1, beta2, scale


def main():
    ref = load_ref()
    rng = np.random.default_rng(20261019)
    out = {}

    # ---- case A: notebook shape (k=50, 250 bins of 0.2 s over -10..40 s), single draw ----------
    k = 50
    edges = np.arange(-10.0, 40.0, 0.2)
    timeA = np.mean([edges[:-1], edges[1:]], axis=0)
    o1, o2, b1, b2, sc = make_params(rng, k, 1, timeA.max() - timeA.min())
    out["A_time"] = timeA
    out["A_omega1"], out["A_omega2"], out["A_beta1"], out["A_beta2"] = o1[:, 0], o2[:, 0], b1[:, 0], b2[:, 0]
    out["A_bw"] = np.float64(0.73)
    out["A_scale"] = sc[:, 0]
    out["A_rff"] = ref.RFF(timeA, o1[:, 0].copy(), o2[:, 0].copy(), b1[:, 0].copy(), b2[:, 0].copy(), 0.73, sc[1, 0])
    out["A_rff_ms"] = ref.RFF_multiscale(
        timeA, o1[:, 0].copy(), o2[:, 0].copy(), b1[:, 0].copy(), b2[:, 0].copy(), 0.73, sc[0, 0], sc[1, 0]
    )

    # ---- case B: draw loop of fit.py:651-691 around the reference RFF; strided (k, S) columns ----
    S = 7
    timeB = np.sort(rng.uniform(-3.0, 25.0, size=401))  # ragged / non-uniform time stamps
    o1, o2, b1, b2, sc = make_params(rng, k, S, 28.0)
    bw = rng.uniform(0.5, 1.5, size=S)
    amp = np.exp(rng.normal(size=S))
    dt = rng.normal(0.0, 0.3, size=S)
    single = np.empty((S, len(timeB)))
    multi = np.empty((S, len(timeB)))
    for n in range(S):  # fit.py:656-666 / 678-689 with the reference's own RFF
        single[n] = amp[n] * ref.RFF(timeB - dt[n], o1[:, n], o2[:, n], b1[:, n], b2[:, n], bw[n], sc[1, n])
        multi[n] = amp[n] * ref.RFF_multiscale(
            timeB - dt[n], o1[:, n], o2[:, n], b1[:, n], b2[:, n], bw[n], sc[0, n], sc[1, n]
        )
    for name, v in dict(time=timeB, omega1=o1, omega2=o2, beta1=b1, beta2=b2, scale=sc, bw=bw, amp=amp, dt=dt,
                        rate_single=single, rate_multi=multi).items():
        out["B_" + name] = v

    # ---- case C: edge cases: one bin, k=1, huge |t| (phase ~ 1e4 rad), zero betas ----------------
    timeC = np.array([1234.5678])
    out["C_time"] = timeC
    out["C_omega1"] = np.array([7.25])
    out["C_omega2"] = np.array([-3.5])
    out["C_beta1"] = np.array([0.31])
    out["C_beta2"] = np.array([-0.77])
    out["C_rff_ms"] = ref.RFF_multiscale(timeC, out["C_omega1"], out["C_omega2"], out["C_beta1"], out["C_beta2"], 1.0, 0.4, 1.3)
    out["C_rff_zero"] = ref.RFF(timeA, np.zeros(3), np.zeros(3), np.zeros(3), np.zeros(3), 1.0, 1.0)  # -> ones

    np.savez_compressed(os.path.join(HERE, "rff_golden.npz"), **out)

    # ---- likelihood composition on a tiny delay grid --------------------------------------------
    lk = {}
    T = 600
    edges = 0.05 * np.arange(T + 1)
    mid = np.mean([edges[:-1], edges[1:]], axis=0)
    expo = edges[1:] - edges[:-1]
    o1, o2, b1, b2, sc = make_params(rng, k, 3, mid.max() - mid.min())
    amp = np.array([40.0, 55.0, 30.0])
    bkg = np.array([500.0, 480.0, 530.0])
    true_dt = 0.4321
    f = ref.RFF_multiscale(mid - true_dt, o1[:, 0], o2[:, 0], b1[:, 0], b2[:, 0], 1.0, sc[0, 0], sc[1, 0])
    counts = rng.poisson(expo * (amp[0] * f + bkg[0])).astype(np.int64)
    counts[5] = 0  # make sure a zero-count bin is present
    delays = np.linspace(-1.0, 1.0, 17)
    ll = np.empty((len(delays), 3))
    for s in range(3):
        for g, d in enumerate(delays):
            fr = ref.RFF_multiscale(mid - d, o1[:, s], o2[:, s], b1[:, s], b2[:, s], 1.0, sc[0, s], sc[1, s])
            ll[g, s] = poisson.logpmf(counts, expo * (amp[s] * fr + bkg[s])).sum()
    for name, v in dict(time=mid, exposure=expo, counts=counts, omega1=o1, omega2=o2, beta1=b1, beta2=b2, scale=sc,
                        amp=amp, bkg=bkg, delays=delays, ll=ll).items():
        lk[name] = v
    np.savez_compressed(os.path.join(HERE, "loglik_golden.npz"), **lk)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    sys.exit(main())
