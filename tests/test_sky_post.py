"""Sky post-processing: host geometry helpers and the oracle's credible-level restatement (CPU),
``ipn_sky_credible`` against the oracle (GPU)."""
import numpy as np
import pytest

from oracle import ipn_oracle as orc


def _constellation(rng, D=5):
    radii = np.array([7.0e3, 4.2e4, 1.5e6, 1.5e8, 2.3e8, 7.0e3, 4.2e4, 1.5e6])[:D]
    v = rng.normal(size=(D, 3))
    return v / np.linalg.norm(v, axis=1, keepdims=True) * radii[:, None]


def test_angles_and_annulus_geometry():
    from pyipn_b200 import sky

    rng = np.random.default_rng(3)
    g = rng.normal(size=(50, 3))
    g /= np.linalg.norm(g, axis=1, keepdims=True)
    phi, theta = sky.grb_phi_theta(g)
    ophi, otheta = orc.grb_angles(g)
    np.testing.assert_allclose(phi, ophi, rtol=0, atol=1e-15)
    np.testing.assert_allclose(theta, otheta, rtol=0, atol=1e-15)
    # the convention is (longitude, latitude): it must invert to the unit vector
    back = np.stack([np.cos(theta) * np.cos(phi), np.cos(theta) * np.sin(phi), np.sin(theta)], axis=1)
    np.testing.assert_allclose(back, g, atol=1e-12)

    pos = _constellation(rng)
    d, n, ra, dec = sky.calculate_distance_and_norm(pos[0], pos[3])
    od, on, ora, odec = orc.calculate_distance_and_norm(pos[0], pos[3])
    assert d == pytest.approx(od, rel=1e-15) and ra == pytest.approx(ora) and dec == pytest.approx(odec)
    np.testing.assert_allclose(n, on, atol=1e-15)
    assert 0.0 <= ra < 2 * np.pi and abs(np.linalg.norm(n) - 1.0) < 1e-15
    # |c dt / d| a hair above 1 (rounding noise) must clip to 0 / pi instead of NaN, like the reference's around(…, 15)
    assert sky.theta_from_time_delay(d / sky.SPEED_OF_LIGHT_KM_S * (1 + 1e-16), d) == 0.0
    assert sky.theta_from_time_delay(0.0, d) == pytest.approx(np.pi / 2)
    with pytest.raises(ValueError):
        sky.calculate_distance_and_norm(pos[0], pos[0])


def test_triangulation_recovers_direction():
    from pyipn_b200 import sky

    rng = np.random.default_rng(11)
    pos = _constellation(rng, D=6)
    g = rng.normal(size=3)
    g /= np.linalg.norm(g)
    # plane wave: arrival time at detector d is -(g . p_d)/c up to a constant
    T0 = -(pos @ g) / sky.SPEED_OF_LIGHT_KM_S
    est = sky.localize_GRB(pos, T0)
    # annulus axis points from detector i to j and cos(theta) = c (T0_i - T0_j)/d = n_ij . g: the solution is g itself
    np.testing.assert_allclose(est, g, atol=1e-9)
    np.testing.assert_allclose(est, orc.localize_grb(pos, T0), atol=1e-12)
    n, radec, theta = sky.calculate_annulus(pos, T0, 0, 1)
    assert np.cos(theta) == pytest.approx(float(n @ g), abs=1e-12)


def test_oracle_credible_levels_properties():
    rng = np.random.default_rng(5)
    P, S = 3000, 7
    ll = rng.normal(size=(P, S)) * 3.0
    ll[:5] = -np.inf           # impossible pixels
    ll[10, :] = ll[11, :]      # an exact tie
    prob, level, order, log_norm = orc.sky_credible(ll)
    assert prob.sum() == pytest.approx(1.0, abs=1e-12) and np.all(prob[:5] == 0.0)
    assert np.all(np.diff(prob[order]) <= 0) and level[order[-1]] == pytest.approx(1.0, abs=1e-12)
    assert list(order).index(10) + 1 == list(order).index(11)  # stable: ties keep pixel order
    # brute-force marginal
    L = np.log(np.mean(np.exp(ll[5:]), axis=1))
    np.testing.assert_allclose(np.log(prob[5:]) + log_norm, L, atol=1e-12)
    # a single draw is the plain softmax
    p1, _, _, z1 = orc.sky_credible(ll[5:, 0])
    np.testing.assert_allclose(p1, np.exp(ll[5:, 0] - z1), rtol=1e-13)


@pytest.mark.gpu
@pytest.mark.parametrize("P,S", [(1, 1), (37, 1), (49152, 1), (49152, 16), (5000, 301)])
def test_sky_credible_vs_oracle(ctx, P, S):
    rng = np.random.default_rng(P + S)
    ll = -0.5 * rng.chisquare(3, size=(P, S)) * 40.0 - 1.0e5
    if P > 30:
        ll[3] = -np.inf
        ll[7] = ll[8]
    lw = rng.normal(size=P) * 0.1 if S > 1 else None
    if lw is not None and P > 30:
        lw[8] = lw[7]
    prob, level, order, log_norm = ctx.sky_credible(ll if S > 1 else ll[:, 0], lw)
    oprob, olevel, oorder, olog_norm = orc.sky_credible(ll, lw)
    assert log_norm == pytest.approx(olog_norm, abs=1e-9)
    np.testing.assert_allclose(prob, oprob, rtol=1e-11, atol=1e-300)
    np.testing.assert_allclose(level, olevel, rtol=0, atol=1e-11)
    assert sorted(order.tolist()) == list(range(P))
    # same order wherever the oracle's neighbours are separated by more than rounding noise
    sp = oprob[oorder]
    clear = np.ones(P, dtype=bool)
    if P > 1:
        gap = np.abs(np.diff(sp)) > 1e-11 * sp[:-1]
        clear[1:] &= gap
        clear[:-1] &= gap
    np.testing.assert_array_equal(order[clear], oorder[clear])
    if P > 30:
        assert prob[3] == 0.0 and list(order).index(7) + 1 == list(order).index(8)


@pytest.mark.gpu
def test_credible_regions_on_healpix_grid(ctx):
    """A Gaussian blob on the nside-32 grid: region areas must match the analytic 2-d Gaussian ones."""
    from pyipn_b200 import sky
    from pyipn_b200.synth import healpix_ring_vectors

    nside = 32
    ghat = healpix_ring_vectors(nside)
    centre = np.array([0.3, -0.5, 0.81])
    centre /= np.linalg.norm(centre)
    sigma = np.deg2rad(6.0)
    ang = np.arccos(np.clip(ghat @ centre, -1, 1))
    ll = -0.5 * (ang / sigma) ** 2
    pix_area = 4 * np.pi / ghat.shape[0]
    out = sky.credible_regions(ll, levels=(0.5, 0.9), pixel_area=pix_area, ctx=ctx)
    assert out["best"] == int(np.argmax(ll))
    for c in (0.5, 0.9):
        analytic = 2 * np.pi * sigma**2 * (-np.log(1 - c))  # flat-sky area of the c-region of a 2-d Gaussian
        assert out["areas"][c] == pytest.approx(analytic, rel=0.08)
        assert out["prob"][out["masks"][c]].sum() >= c
    with pytest.raises(ValueError):
        sky.credible_regions(ll, levels=(1.5,), ctx=ctx)
    with pytest.raises(Exception):
        ctx.sky_credible(np.full(10, -np.inf))
