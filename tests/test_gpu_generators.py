"""GPU, next-row N2: thinning generators (pyipn/possion_gen.py:159-242) -- statistical parity with the host generator that
replays the reference's MT19937 stream, determinism, and the full device data path events -> bins -> likelihood."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _chi2(obs, exp):
    m = exp > 5
    return ((obs[m] - exp[m]) ** 2 / exp[m]).sum(), int(m.sum())


def test_thinning_matches_intensity_and_host_generator(ctx):
    from pyipn_b200.possion_gen import (background_poisson_generator, background_poisson_generator_gpu, norris,
                                        source_poisson_generator, source_poisson_generator_gpu)

    tstart, tstop = -5.0, 30.0
    src = source_poisson_generator_gpu(tstart, tstop, 3000.0, 0.0, 0.5, 4.0, seed=11)
    assert src[0] == tstart and np.all(np.diff(src[1:]) >= 0) and np.all(src[1:] > 0.0) and src[-1] <= tstop
    edges = np.linspace(tstart, tstop, 141)
    mid = 0.5 * (edges[1:] + edges[:-1])
    fine = np.linspace(tstart, tstop, 140 * 50 + 1)
    expected = np.add.reduceat(norris(0.5 * (fine[1:] + fine[:-1]), 3000.0, 0.0, 0.5, 4.0) * np.diff(fine), np.arange(0, 7000, 50))
    obs = np.histogram(src[1:], bins=edges)[0]
    c2, dof = _chi2(obs, expected)
    assert c2 < dof + 5 * np.sqrt(2 * dof), (c2, dof)
    # two-sample test against the host generator (reference stream)
    host = np.histogram(source_poisson_generator(tstart, tstop, 3000.0, 0.0, 0.5, 4.0, seed=11)[1:], bins=edges)[0]
    m = (obs + host) > 10
    c2 = ((obs[m] - host[m]) ** 2 / (obs[m] + host[m])).sum()
    assert c2 < m.sum() + 5 * np.sqrt(2 * m.sum())
    # determinism and seed dependence
    np.testing.assert_array_equal(src, source_poisson_generator_gpu(tstart, tstop, 3000.0, 0.0, 0.5, 4.0, seed=11))
    assert len(src) != len(source_poisson_generator_gpu(tstart, tstop, 3000.0, 0.0, 0.5, 4.0, seed=12)) or not np.array_equal(
        src, source_poisson_generator_gpu(tstart, tstop, 3000.0, 0.0, 0.5, 4.0, seed=12))
    # background: flat 500 cps and a sloped one
    bkg = background_poisson_generator_gpu(tstart, tstop, 0.0, 500.0, seed=3)
    n = len(bkg) - 1
    assert abs(n - 500 * 35) < 5 * np.sqrt(500 * 35)
    host_n = len(background_poisson_generator(tstart, tstop, 0.0, 500.0, seed=3)) - 1
    assert abs(n - host_n) < 6 * np.sqrt(2 * 500 * 35)
    sl = background_poisson_generator_gpu(0.0, 10.0, 20.0, 100.0, seed=5)[1:]
    h = np.histogram(sl, bins=10, range=(0, 10))[0]
    exp = 100.0 + 20.0 * (np.arange(10) + 0.5)
    assert _chi2(h, exp)[0] < 10 + 5 * np.sqrt(20)
    # no source photons expected -> empty array, like the reference (possion_gen.py:188-190)
    assert len(source_poisson_generator_gpu(0.0, 5.0, 300.0, 100.0, 0.5, 4.0, seed=1)) == 0
    # multi-pulse
    mp = source_poisson_generator_gpu(0.0, 20.0, [1000.0, 500.0], [1.0, 8.0], [0.3, 0.5], [1.0, 2.0], seed=9)
    assert np.all(mp[1:] > 1.0) and np.histogram(mp[1:], bins=[0, 8, 20])[0].min() > 100


def test_device_data_path_events_to_likelihood(ctx):
    """events (GPU thinning) -> counts (GPU histogram) -> delay-grid likelihood (GPU), recovered delay."""
    from pyipn_b200 import LightCurve
    from pyipn_b200.possion_gen import background_poisson_generator_gpu, source_poisson_generator_gpu

    lag = 0.4
    lcs = []
    for d, t0 in enumerate((0.0, lag)):
        s = source_poisson_generator_gpu(-5.0, 25.0, 4000.0, t0, 0.5, 4.0, seed=100 + 10 * d)
        b = background_poisson_generator_gpu(-5.0, 25.0, 0.0, 500.0, seed=101 + 10 * d)
        lcs.append(LightCurve(s, b))
    _, e0, c0 = lcs[0].get_binned_light_curve(-4.0, 24.0, 0.01, device=True)
    _, e1, c1 = lcs[1].get_binned_light_curve(-4.0, 24.0, 0.01, device=True)
    # template: Fourier fit (the RFF functional form with fixed frequencies) of detector 0's smoothed total log-rate
    mid = 0.5 * (e0[1:] + e0[:-1])
    k = 50
    span = mid[-1] - mid[0]
    omega = 2 * np.pi * np.arange(1, k + 1) / (1.5 * span)
    sm = np.convolve(np.pad(c0, 25, mode="edge"), np.ones(51) / 51, mode="valid") / 0.01
    logf = np.log(np.maximum(sm, 50.0))
    A = np.concatenate([np.cos(np.outer(mid, omega)), np.sin(np.outer(mid, omega))], axis=1)
    coef = np.linalg.lstsq(A, logf - logf.mean(), rcond=1e-3)[0]  # truncated SVD: the non-periodic window is ill-conditioned
    assert np.abs(coef).max() < 5.0
    w = np.concatenate([omega, omega])[None, :]
    a = np.concatenate([coef[:k], np.zeros(k)])[None, :]
    b = np.concatenate([np.zeros(k), coef[k:]])[None, :]
    amp = np.array([np.exp(logf.mean())])
    delays = np.linspace(-1.0, 1.0, 201)
    core = slice(300, -300)  # keep away from the window edges (the template is periodic)
    ll0 = ctx.loglik_delay_grid(mid[core], np.diff(e0)[core], c0[core], w, a, b, amp, [1e-3], delays)
    ll1 = ctx.loglik_delay_grid(mid[core], np.diff(e1)[core], c1[core], w, a, b, amp, [1e-3], delays)
    assert abs(delays[int(np.argmax(ll0[:, 0]))]) < 0.06
    assert abs(delays[int(np.argmax(ll1[:, 0]))] - lag) < 0.06
