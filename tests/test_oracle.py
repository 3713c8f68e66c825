"""CPU: pin the oracle (oracle/ipn_oracle.py) against the fixtures minted from the reference's own rff.py."""
import numpy as np
import pytest
from scipy.stats import poisson

from oracle import ipn_oracle as orc


def test_rff_matches_reference_golden(rff_golden):
    g = rff_golden
    got = orc.rff(g["A_time"], g["A_omega1"], g["A_omega2"], g["A_beta1"], g["A_beta2"], float(g["A_bw"]), g["A_scale"][1])
    np.testing.assert_allclose(got, g["A_rff"], rtol=1e-12)
    got = orc.rff_multiscale(g["A_time"], g["A_omega1"], g["A_omega2"], g["A_beta1"], g["A_beta2"], float(g["A_bw"]),
                             g["A_scale"][0], g["A_scale"][1])
    np.testing.assert_allclose(got, g["A_rff_ms"], rtol=1e-12)


def test_expected_rate_loop_matches_reference_golden(rff_golden):
    g = rff_golden
    args = (g["B_time"], g["B_omega1"], g["B_omega2"], g["B_beta1"], g["B_beta2"], g["B_bw"])
    single = orc.expected_rate(*args, g["B_scale"][1], g["B_amp"], g["B_dt"], 7)
    multi = orc.expected_rate_multiscale(*args, g["B_scale"], g["B_amp"], g["B_dt"], 7)
    np.testing.assert_allclose(single, g["B_rate_single"], rtol=1e-11)
    np.testing.assert_allclose(multi, g["B_rate_multi"], rtol=1e-11)


def test_rff_edge_cases_golden(rff_golden):
    g = rff_golden
    got = orc.rff_multiscale(g["C_time"], g["C_omega1"], g["C_omega2"], g["C_beta1"], g["C_beta2"], 1.0, 0.4, 1.3)
    np.testing.assert_allclose(got, g["C_rff_ms"], rtol=1e-10)  # |phase| ~ 9e3 rad
    np.testing.assert_array_equal(g["C_rff_zero"], np.ones_like(g["C_rff_zero"]))
    assert orc.rff(np.zeros(0), np.ones(2), np.ones(2), np.ones(2), np.ones(2), 1.0, 1.0).shape == (0,)


def test_folded_form_equals_reference_form(rff_golden):
    g = rff_golden
    for s in range(7):
        w, a, b = orc.fold_params(g["B_omega1"][:, s], g["B_omega2"][:, s], g["B_beta1"][:, s], g["B_beta2"][:, s],
                                  g["B_bw"][s], g["B_scale"][0, s], g["B_scale"][1, s])
        lam = g["B_amp"][s] * np.exp(orc.log_intensity_folded(g["B_time"] - g["B_dt"][s], w, a, b))
        np.testing.assert_allclose(lam, g["B_rate_multi"][s], rtol=1e-11)


def test_poisson_lpmf_matches_scipy_and_stan_edge_cases():
    rng = np.random.default_rng(1)
    mu = rng.uniform(0.01, 30.0, 500)
    n = rng.poisson(mu)
    assert abs(orc.poisson_lpmf(n, mu) - poisson.logpmf(n, mu).sum()) < 1e-9
    assert orc.poisson_lpmf([0, 0], [0.0, 0.0]) == 0.0
    assert orc.poisson_lpmf([1], [0.0]) == -np.inf
    assert abs(orc.poisson_lpmf([0], [2.5]) + 2.5) < 1e-15


def test_loglik_composition_matches_reference_golden(loglik_golden):
    g = loglik_golden
    for s in range(3):
        for gi, d in enumerate(g["delays"]):
            got = orc.loglik_bins(g["counts"], g["time"], g["exposure"], g["omega1"][:, s], g["omega2"][:, s],
                                  g["beta1"][:, s], g["beta2"][:, s], d, g["bkg"][s], g["scale"][0, s], g["scale"][1, s],
                                  g["amp"][s])
            assert abs(got - g["ll"][gi, s]) < 1e-7
    w, a, b = zip(*[orc.fold_params(g["omega1"][:, s], g["omega2"][:, s], g["beta1"][:, s], g["beta2"][:, s], 1.0,
                                    g["scale"][0, s], g["scale"][1, s]) for s in range(3)])
    ll = orc.loglik_delay_grid(g["counts"], g["time"], g["exposure"], np.array(w), np.array(a), np.array(b), g["amp"],
                               g["bkg"], g["delays"])
    np.testing.assert_allclose(ll, g["ll"], atol=1e-7, rtol=0)
    assert np.argmax(ll[:, 0]) == np.argmax(g["ll"][:, 0])


def test_total_loglik_ragged_detectors(loglik_golden):
    g = loglik_golden
    T = len(g["time"])
    counts = np.zeros((2, T), dtype=int)
    time = np.zeros((2, T))
    expo = np.zeros((2, T))
    nb = [T, T - 100]
    for d in range(2):
        counts[d, : nb[d]] = g["counts"][: nb[d]]
        time[d, : nb[d]] = g["time"][: nb[d]]
        expo[d, : nb[d]] = g["exposure"][: nb[d]]
    args = (g["omega1"][:, 0], g["omega2"][:, 0], g["beta1"][:, 0], g["beta2"][:, 0])
    tot = orc.total_loglik(counts, time, expo, nb, *args, [0.0, 0.3], [500.0, 480.0], g["scale"][0, 0], g["scale"][1, 0], [1.0, 0.9])
    parts = sum(orc.loglik_bins(counts[d][: nb[d]], time[d][: nb[d]], expo[d][: nb[d]], *args, [0.0, 0.3][d],
                                [500.0, 480.0][d], g["scale"][0, 0], g["scale"][1, 0], [1.0, 0.9][d]) for d in range(2))
    assert tot == pytest.approx(parts, abs=1e-9)


def test_time_delay_and_sky_delays():
    sc = np.array([[7000.0, 0, 0], [0.0, 42000.0, 0], [1.5e6, 2.0e5, -3.0e4]])
    g = np.array([0.6, 0.0, 0.8])
    d = orc.sky_delays(g, sc)[0]
    assert d[0] == 0.0
    assert d[1] == pytest.approx(np.dot(g, sc[0] - sc[1]) / 299792.46, rel=1e-15)
    assert orc.time_delay(g, sc[0] - sc[2]) == pytest.approx(d[2], rel=1e-15)
    phi, theta = orc.grb_angles(g)
    assert phi == pytest.approx(0.0) and theta == pytest.approx(np.arcsin(0.8))


def test_healpix_ring_properties():
    from pyipn_b200.synth import healpix_ring_vectors

    for nside in (1, 2, 4, 8):
        npix = 12 * nside * nside
        v = orc.healpix_ring_pix2vec(nside, np.arange(npix))
        np.testing.assert_allclose(np.linalg.norm(v, axis=1), 1.0, atol=1e-14)
        np.testing.assert_allclose(v.sum(axis=0), 0.0, atol=1e-10)  # pixel centres are symmetric
        assert np.all(np.diff(v[:, 2]) <= 1e-15)  # RING order: z never increases
        assert len(np.unique(np.round(v[:, 2], 12))) == 4 * nside - 1  # number of iso-latitude rings
        np.testing.assert_allclose(healpix_ring_vectors(nside), v, atol=1e-14)  # product-side vectorised twin
    # known value: nside=1 pixel 0 is at z = 2/3, phi = pi/4
    v = orc.healpix_ring_pix2vec(1, [0])[0]
    assert v[2] == pytest.approx(2.0 / 3.0) and np.arctan2(v[1], v[0]) == pytest.approx(np.pi / 4)


def test_binning_and_stan_data_layout():
    rng = np.random.default_rng(3)
    ev = [np.sort(rng.uniform(-5, 30, 4000)), np.sort(rng.uniform(-5, 30, 3000))]
    rate, edges, counts = orc.bin_events(ev[0], -2.0, 20.0, 0.2)
    assert len(edges) == len(np.arange(-2.0, 20.0, 0.2)) and len(counts) == len(edges) - 1
    np.testing.assert_allclose(rate, counts / 0.2)
    d = orc.to_stan_data(ev, np.eye(3)[:2] * 7000.0, [-2.0, 0.0], [20.0, 15.0], [0.2, 0.1], k=50)
    assert d["N_time_bins"] == [len(np.arange(-2, 20, 0.2)) - 1, len(np.arange(0, 15, 0.1)) - 1]
    assert d["counts"].shape == (2, max(d["N_time_bins"]))
    assert d["counts"][0, d["N_time_bins"][0]:].sum() == 0 and d["exposure"][0, d["N_time_bins"][0]:].sum() == 0
    np.testing.assert_allclose(d["exposure"][1, : d["N_time_bins"][1]], 0.1)


def test_draw_prior_respects_constraints():
    rng = np.random.default_rng(5)
    for _ in range(20):
        d = orc.draw_prior(rng, 50, 100.0, n_detectors=3)
        assert 0 < d["scale"][0] <= d["scale"][1]
        assert d["bw"][0] < d["bw"][1]  # range2 < range1  ->  bw2 > bw1
        assert d["omega"].shape == (2, 50) and d["amplitude"].shape == (2,) and d["bkg"].shape == (3,)


def test_ppc_counts_stream_is_numpy_legacy():
    rates = np.full((2, 5), 3.0)
    out = orc.ppc_counts(np.arange(5.0), np.full(5, 0.5), rates, np.array([1.0, 2.0]), seed=7)
    rs = np.random.RandomState(7)
    exp = np.array([[rs.poisson((3.0 + b) * 0.5) for _ in range(5)] for b in (1.0, 2.0)], dtype=float)
    np.testing.assert_array_equal(out, exp)
