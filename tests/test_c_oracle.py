"""CPU: pin the plain-C oracle (oracle/ipn_oracle.c) against the fixtures minted from the reference's own rff.py and
against the numpy restatement."""
import numpy as np
import pytest

from oracle import c_oracle as corc
from oracle import ipn_oracle as orc


@pytest.fixture(scope="module", autouse=True)
def _built():
    corc.build()


def test_c_expected_rate_matches_reference_golden(rff_golden):
    g = rff_golden
    args = (g["B_time"], g["B_omega1"], g["B_omega2"], g["B_beta1"], g["B_beta2"], g["B_bw"])
    single = corc.expected_rate(*args, g["B_scale"][1], g["B_amp"], g["B_dt"], 7)
    multi = corc.expected_rate(*args, g["B_scale"], g["B_amp"], g["B_dt"], 7)
    np.testing.assert_allclose(single, g["B_rate_single"], rtol=1e-11)
    np.testing.assert_allclose(multi, g["B_rate_multi"], rtol=1e-11)
    # single evaluation A (k,) arrays are the S = 1 case of the (k, S) layout
    one = corc.expected_rate(g["A_time"], g["A_omega1"][:, None], g["A_omega2"][:, None], g["A_beta1"][:, None],
                             g["A_beta2"][:, None], [float(g["A_bw"])], g["A_scale"][:, None], [1.0], [0.0])
    np.testing.assert_allclose(one[0], g["A_rff_ms"], rtol=1e-12)
    # fewer draws than columns (N < S) and the |phase| ~ 9e3 rad edge case
    np.testing.assert_allclose(corc.expected_rate(*args, g["B_scale"], g["B_amp"], g["B_dt"], 3), g["B_rate_multi"][:3], rtol=1e-11)
    edge = corc.expected_rate(g["C_time"], g["C_omega1"][:, None], g["C_omega2"][:, None], g["C_beta1"][:, None],
                              g["C_beta2"][:, None], [1.0], np.array([[0.4], [1.3]]), [1.0], [0.0])
    np.testing.assert_allclose(edge[0], g["C_rff_ms"], rtol=1e-10)


def test_c_loglik_matches_reference_golden(loglik_golden):
    g = loglik_golden
    w, a, b = zip(*[orc.fold_params(g["omega1"][:, s], g["omega2"][:, s], g["beta1"][:, s], g["beta2"][:, s], 1.0,
                                    g["scale"][0, s], g["scale"][1, s]) for s in range(3)])
    ll = corc.loglik_delay_grid(g["counts"], g["time"], g["exposure"], np.array(w), np.array(a), np.array(b), g["amp"],
                                g["bkg"], g["delays"])
    np.testing.assert_allclose(ll, g["ll"], atol=1e-7, rtol=0)
    assert np.argmax(ll[:, 0]) == np.argmax(g["ll"][:, 0])


@pytest.mark.parametrize("threads", [1, 3, 0])
def test_c_oracle_equals_numpy_oracle(threads):
    from pyipn_b200.synth import delay_grid_workload, sky_grid_workload

    corc.set_threads(threads)
    try:
        w = delay_grid_workload(T=3000, G=9, S=2)
        args = (w["counts"], w["time"], w["exposure"], w["omega"], w["a"], w["b"], w["amp"], w["bkg"], w["delays"])
        np.testing.assert_allclose(corc.loglik_delay_grid(*args), orc.loglik_delay_grid(*args), rtol=1e-13, atol=1e-9)
        # zero counts, zero amplitude (mu = exposure * bkg), an empty light curve
        z = list(args)
        z[0] = np.zeros_like(w["counts"])
        z[6] = np.zeros(2)
        np.testing.assert_allclose(corc.loglik_delay_grid(*z), orc.loglik_delay_grid(*z), rtol=1e-13, atol=1e-9)
        e = [x[:0] if i < 3 else x for i, x in enumerate(args)]
        assert np.all(corc.loglik_delay_grid(*e) == 0.0)
        s = sky_grid_workload(D=3, T=400, S=2, nside=1, dt_bin=0.02)
        s["nbins"][1] = 333  # ragged
        sargs = (s["counts"], s["time"], s["exposure"], s["nbins"], s["sc_pos"], s["ghat"], s["omega"], s["a"], s["b"], s["amp"],
                 s["bkg"])
        np.testing.assert_allclose(corc.loglik_sky_grid(*sargs), orc.loglik_sky_grid(*sargs), rtol=1e-13, atol=1e-9)
    finally:
        corc.set_threads(0)


def test_c_oracle_random_shapes_property():
    """Hypothesis: for random small shapes (including empty axes and zero exposure) the C and numpy restatements agree."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    @settings(max_examples=40, deadline=None, derandomize=True, database=None)
    @given(T=st.integers(0, 40), G=st.integers(1, 5), S=st.integers(1, 3), J=st.integers(1, 9), seed=st.integers(0, 2**31 - 1))
    def check(T, G, S, J, seed):
        rng = np.random.default_rng(seed)
        time = np.sort(rng.uniform(-5, 5, T))
        expo = rng.uniform(0.0, 0.3, T)
        if T > 2:
            expo[1] = 0.0  # mu = 0 with n = 0 must contribute nothing
        omega = rng.normal(size=(S, J)) * 3.0
        a = rng.normal(size=(S, J)) * 0.3
        b = rng.normal(size=(S, J)) * 0.3
        amp = np.exp(rng.normal(2.0, 1.0, S))
        bkg = np.exp(rng.normal(1.0, 1.0, S))
        counts = rng.poisson(expo * 20.0).astype(np.int64)
        delays = rng.uniform(-1, 1, G)
        got = corc.loglik_delay_grid(counts, time, expo, omega, a, b, amp, bkg, delays)
        exp = orc.loglik_delay_grid(counts, time, expo, omega, a, b, amp, bkg, delays)
        np.testing.assert_allclose(got, exp, rtol=1e-12, atol=1e-10)

    check()
