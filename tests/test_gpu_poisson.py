"""GPU: ipn_poisson_counts (drop-in for ppc_generator, pyipn/fit.py:694-705) -- statistical parity with the
MT19937 oracle stream (a parallel counter-based generator cannot be stream-identical)."""
import numpy as np
import pytest

from oracle import ipn_oracle as orc

pytestmark = pytest.mark.gpu


def test_ppc_generator_moments_and_determinism(ctx):
    from pyipn_b200 import ppc_generator

    T, S = 4000, 6
    rates = np.tile(np.array([0.0, 0.5, 3.0, 9.99, 10.0, 250.0])[:, None] / 0.5, (1, T))
    bkg = np.zeros(S)
    expo = np.full(T, 0.5)
    out = ppc_generator(np.arange(T), expo, rates, bkg, S, seed=11)
    assert out.shape == (S, T) and out.dtype == np.float64
    assert np.all(out == np.floor(out)) and np.all(out >= 0)
    lam = rates[:, 0] * 0.5
    np.testing.assert_array_equal(out[0], 0.0)
    for s in range(1, S):
        assert abs(out[s].mean() - lam[s]) < 5 * np.sqrt(lam[s] / T)
        assert abs(out[s].var() - lam[s]) < 0.12 * lam[s]
    np.testing.assert_array_equal(out, ppc_generator(np.arange(T), expo, rates, bkg, S, seed=11))
    assert np.any(out != ppc_generator(np.arange(T), expo, rates, bkg, S, seed=12))


def test_ppc_distribution_matches_oracle_stream(ctx):
    """Two-sample comparison against the reference-stream oracle (numpy legacy MT19937 == numba's stream)."""
    from pyipn_b200 import ppc_generator

    T = 3000
    rates = np.full((1, T), 6.0)
    ours = ppc_generator(np.arange(T), np.ones(T), rates, np.array([1.5]), 1, seed=3)[0]
    ref = orc.ppc_counts(np.arange(T), np.ones(T), rates, np.array([1.5]), seed=3)[0]
    hist_o = np.bincount(ours.astype(int), minlength=30)[:30]
    hist_r = np.bincount(ref.astype(int), minlength=30)[:30]
    m = (hist_o + hist_r) > 10
    chi2 = ((hist_o[m] - hist_r[m]) ** 2 / (hist_o[m] + hist_r[m])).sum()
    assert chi2 < 3.0 * m.sum()
