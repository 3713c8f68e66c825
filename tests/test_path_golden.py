"""Oracle, host mirror and CUDA path against ``tests/golden/path_golden.npz``: outputs of the reference's OWN functions
either side of the RFF intensity (pyipn/fit.py:651-705 draw loops and ppc_generator, pyipn/possion_gen.py:159-242 thinning
generators, pyipn/lightcurve.py:24-47 binning, pyipn/universe.py:441-529 to_stan_data), lifted out of the reference files
and run in the build container by ``tests/golden/make_golden_path.py``."""
import collections
import os

import numpy as np
import pytest

from oracle import ipn_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "path_golden.npz")


@pytest.fixture(scope="module")
def g():
    return dict(np.load(GOLDEN))


def _loop_args(g):
    return g["F_time"], g["F_omega1"], g["F_omega2"], g["F_beta1"], g["F_beta2"], g["F_bw"]


# ---- oracle (numpy and C restatements) -------------------------------------------------------------------------------
def test_oracle_draw_loops_match_reference_functions(g):
    S = len(g["F_amp"])
    single = orc.expected_rate(*_loop_args(g), g["F_scale"][1], g["F_amp"], g["F_dt"], S)
    multi = orc.expected_rate_multiscale(*_loop_args(g), g["F_scale"], g["F_amp"], g["F_dt"], S)
    # the reference compiles the loops with fastmath=True: its own result is only defined to a few ulp of the sum
    np.testing.assert_allclose(single, g["F_rate_single"], rtol=1e-11)
    np.testing.assert_allclose(multi, g["F_rate_multi"], rtol=1e-11)
    # N < S evaluates the first N draws only (fit.py:654-656)
    np.testing.assert_allclose(orc.expected_rate_multiscale(*_loop_args(g), g["F_scale"], g["F_amp"], g["F_dt"], 3),
                               g["F_rate_multi"][:3], rtol=1e-11)


def test_c_oracle_draw_loops_match_reference_functions(g):
    from oracle import c_oracle

    c_oracle.build()
    S = len(g["F_amp"])
    np.testing.assert_allclose(c_oracle.expected_rate(*_loop_args(g), g["F_scale"][1], g["F_amp"], g["F_dt"], S),
                               g["F_rate_single"], rtol=1e-11)
    np.testing.assert_allclose(c_oracle.expected_rate(*_loop_args(g), g["F_scale"], g["F_amp"], g["F_dt"], S),
                               g["F_rate_multi"], rtol=1e-11)


def test_oracle_ppc_stream_matches_reference_ppc_generator(g):
    # the reference under numba (MT19937 seeded from compiled code) and as plain Python agree; the oracle reproduces both
    np.testing.assert_array_equal(g["P_counts_numba"], g["P_counts_numpy"])
    out = orc.ppc_counts(np.zeros(60), g["P_exposure"], g["P_rates"], g["P_bkg"], int(g["P_seed"]))
    np.testing.assert_array_equal(out, g["P_counts_numba"])
    assert np.all(out[3][g["P_rates"][3] * g["P_exposure"] == 0.0] == 0.0)


def test_oracle_binning_and_stan_data_match_reference(g):
    events = np.concatenate([g["G_src"], g["G_bkg"]])
    rate, edges, counts = orc.bin_events(events, *g["L_window"])
    np.testing.assert_array_equal(counts, g["L_counts"])
    np.testing.assert_array_equal(edges, g["L_edges"])
    np.testing.assert_array_equal(rate, g["L_rate"])
    per_det = [events, np.concatenate([g["G_src_multi"], g["G_bkg_slope"]])]
    sd = orc.to_stan_data(per_det, g["D_sc_pos"], [-1.0, 0.5], [6.0, 3.75], [0.2, 0.05], k=25, n_cores=2, factor=0.5, factor2=1)
    for key in ("N_detectors", "N_time_bins", "max_N_time_bins", "counts", "time", "exposure", "sc_pos", "k", "grainsize",
                "grainsize_meta"):
        np.testing.assert_array_equal(np.asarray(sd[key]), g["D2_" + key], err_msg=key)
    sd1 = orc.to_stan_data(per_det, g["D_sc_pos"], -1.0, 6.0, 0.2)
    for key in ("N_time_bins", "counts", "time", "exposure", "grainsize", "grainsize_meta"):
        np.testing.assert_array_equal(np.asarray(sd1[key]), g["D1_" + key], err_msg=key)


def test_reference_single_selection_slip_is_what_the_docs_say(g):
    """universe.py:466-468 reuses the loop variable for "fall back to the first time selection": with one selection and two
    detectors, detector 1's position lands in row 0 and row 1 is never written.  The mirror keeps the binning behaviour
    (checked above) and does not reproduce the misplaced rows (pyipn_b200/universe.py)."""
    np.testing.assert_array_equal(g["D1_sc_pos"][0], g["D_sc_pos"][1])
    np.testing.assert_array_equal(g["D1_effective_area"][0], g["D_area"][1])
    np.testing.assert_array_equal(g["D2_sc_pos"], g["D_sc_pos"])  # full-length selections: rows where they belong


# ---- host mirror (pyipn_b200 facade, CPU parts) ----------------------------------------------------------------------
def test_host_generators_reproduce_reference_arrival_times(g):
    from pyipn_b200.possion_gen import background_poisson_generator, pulse, source_poisson_generator

    a = g["G_src_args"]
    src = source_poisson_generator(a[0], a[1], a[2], a[3], a[4], a[5], seed=int(a[6]))
    assert src.shape == g["G_src"].shape  # same accept/reject decisions, draw for draw
    np.testing.assert_allclose(src, g["G_src"], rtol=0, atol=1e-9)
    a = g["G_bkg_args"]
    bkg = background_poisson_generator(a[0], a[1], a[2], a[3], seed=int(a[4]))
    assert bkg.shape == g["G_bkg"].shape
    np.testing.assert_allclose(bkg, g["G_bkg"], rtol=0, atol=1e-9)
    a = g["G_bkg_slope_args"]
    bkg = background_poisson_generator(a[0], a[1], a[2], a[3], seed=int(a[4]))
    assert bkg.shape == g["G_bkg_slope"].shape
    np.testing.assert_allclose(bkg, g["G_bkg_slope"], rtol=0, atol=1e-9)
    Ks, ts, tr, td = g["G_src_multi_pulses"]
    multi = source_poisson_generator(-1.0, 7.0, Ks, ts, tr, td, seed=99)
    assert multi.shape == g["G_src_multi"].shape
    np.testing.assert_allclose(multi, g["G_src_multi"], rtol=0, atol=1e-9)
    assert len(source_poisson_generator(0.0, 5.0, 500.0, 100.0, 0.5, 4.1, seed=1)) == len(g["G_src_late"]) == 0
    np.testing.assert_allclose(pulse(g["G_pulse_x"], 500.0, 0.0, 0.5, 4.1), g["G_pulse"], rtol=1e-13)
    np.testing.assert_allclose(pulse(g["G_pulse_x"], Ks, ts, tr, td), g["G_pulse_multi"], rtol=1e-13)


class _Det:
    def __init__(self, xyz, pointing, area):
        self.xyz_km, self.pointing_vec, self._area = xyz, pointing, area

    def effective_area_towards(self, grb):
        return self._area


def _universe(g):
    from pyipn_b200 import Universe
    from pyipn_b200.lightcurve import LightCurve

    uni = Universe(grb=None)
    for i in range(2):
        d = _Det(g["D_sc_pos"][i], g["D_pointing"][i], g["D_area"][i])
        d.name = f"det{i}"
        uni.register_detector(d)
    uni.set_light_curves(collections.OrderedDict(det0=LightCurve(g["G_src"], g["G_bkg"]),
                                                 det1=LightCurve(g["G_src_multi"], g["G_bkg_slope"])))
    return uni


def test_host_lightcurve_and_universe_match_reference(g):
    uni = _universe(g)
    rate, edges, counts = uni.light_curves["det0"].get_binned_light_curve(*g["L_window"])
    np.testing.assert_array_equal(counts, g["L_counts"])
    np.testing.assert_array_equal(edges, g["L_edges"])
    np.testing.assert_array_equal(rate, g["L_rate"])
    sd = uni.to_stan_data([-1.0, 0.5], [6.0, 3.75], [0.2, 0.05], k=25, n_cores=2, factor=0.5, factor2=1)
    assert set(sd) == {k[3:] for k in g if k.startswith("D2_")}  # same keys as the reference's dict
    for key in sd:
        np.testing.assert_array_equal(np.asarray(sd[key]), g["D2_" + key], err_msg=key)
        assert np.asarray(sd[key]).dtype.kind == g["D2_" + key].dtype.kind, key
    sd1 = uni.to_stan_data(-1.0, 6.0, 0.2)
    for key in ("N_detectors", "N_time_bins", "max_N_time_bins", "counts", "time", "exposure", "k", "grainsize", "grainsize_meta"):
        np.testing.assert_array_equal(np.asarray(sd1[key]), g["D1_" + key], err_msg=key)


# ---- CUDA path through the C ABI -------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_draw_loops_match_reference_functions(ctx, g):
    from pyipn_b200.fit import _expeced_rate, _expeced_rate_multiscale

    S = len(g["F_amp"])
    single = _expeced_rate(*_loop_args(g), g["F_scale"][1], g["F_amp"], g["F_dt"], S)
    multi = _expeced_rate_multiscale(*_loop_args(g), g["F_scale"], g["F_amp"], g["F_dt"], S)
    assert single.shape == multi.shape == (S, len(g["F_time"])) and single.dtype == np.float64
    np.testing.assert_allclose(single, g["F_rate_single"], rtol=1e-6)  # north-star tolerance on lambda
    np.testing.assert_allclose(multi, g["F_rate_multi"], rtol=1e-6)
    np.testing.assert_allclose(_expeced_rate_multiscale(*_loop_args(g), g["F_scale"], g["F_amp"], g["F_dt"], 4),
                               g["F_rate_multi"][:4], rtol=1e-6)


@pytest.mark.gpu
def test_gpu_binning_and_stan_data_match_reference(ctx, g):
    uni = _universe(g)
    rate, edges, counts = uni.light_curves["det0"].get_binned_light_curve(*g["L_window"], device=True)
    np.testing.assert_array_equal(counts, g["L_counts"])  # integer work: bit-exact
    np.testing.assert_array_equal(edges, g["L_edges"])
    np.testing.assert_array_equal(rate, g["L_rate"])
    ev1 = np.concatenate([g["G_src_multi"], g["G_bkg_slope"]])
    _, e1, c1 = ctx.bin_events(ev1, 0.5, 3.75, 0.05)
    n1 = int(g["D2_N_time_bins"][1])
    np.testing.assert_array_equal(c1, g["D2_counts"][1, :n1])
    np.testing.assert_array_equal(e1[1:] - e1[:-1], g["D2_exposure"][1, :n1])


@pytest.mark.gpu
def test_gpu_ppc_generator_matches_reference_moments(ctx, g):
    """Philox stream, so parity with the reference's MT19937 counts is statistical: same Poisson means."""
    from pyipn_b200.fit import ppc_generator

    mu = (g["P_rates"] + g["P_bkg"][:, None]) * g["P_exposure"][None, :]
    reps = np.stack([ppc_generator(np.zeros(60), g["P_exposure"], g["P_rates"], g["P_bkg"], 4, seed=s) for s in range(64)])
    assert reps.shape == (64, 4, 60) and reps.dtype == np.float64 and np.all(reps == np.rint(reps)) and np.all(reps >= 0)
    z = (reps.mean(axis=0) - mu) / np.sqrt(np.maximum(mu, 1e-12) / 64)
    assert np.all(np.abs(z[mu > 0]) < 5.5)
    assert np.all(reps[:, mu == 0] == 0)
    zref = (g["P_counts_numba"] - mu)[mu > 0] / np.sqrt(mu[mu > 0])  # the reference's own draw sits in the same distribution
    assert np.all(np.abs(zref) < 5.5)
