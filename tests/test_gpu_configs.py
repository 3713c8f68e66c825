"""GPU: the BASELINE.json configurations end to end through the public facade, each checked against the oracle.

config 1  examples/template_3.yaml-like universe (2 active detectors, K=500, t_rise=0.5, t_decay=4.1) -> to_stan_data
          (0.2 s bins, k=50) -> delay grid and sky grid of both detectors
config 4  8-detector constellation, HEALPix nside=64 (49152 pixels), 1e5 bins per detector: full size, oracle on a
          pixel subsample, exact argmax among the checked pixels, summed log-likelihood per pixel
config 5  population run (reduced to 3 GRBs x 3 detectors here): RFF intensity -> Poisson counts on the GPU ->
          delay-grid likelihood of every detector against detector 0; recovered delays
(configs 2 and 3 are covered in test_gpu_loglik.py and by bench.py)"""
import numpy as np
import pytest

from oracle import ipn_oracle as orc

pytestmark = pytest.mark.gpu
ATOL = 1e-3

TEMPLATE_3 = dict(
    grb=dict(ra=180.0, dec=30.0, distance=500.0, K=500.0, t_rise=0.5, t_decay=4.1),
    detectors=dict(
        det3=dict(ra=300.0, dec=-45.0, altitude=153000.0, time="2010-01-01T00:00:00", pointing=dict(ra=180.0, dec=30.0),
                  effective_area=1.0),
        det2=dict(ra=60.0, dec=10.0, altitude=500.0, time="2010-01-01T00:00:00", pointing=dict(ra=180.0, dec=30.0),
                  effective_area=1.0),
    ),
)


def test_config1_template3_universe_to_likelihood(ctx):
    from pyipn_b200 import Fit, Universe
    from pyipn_b200.synth import draw_prior

    uni = Universe.from_dict(TEMPLATE_3)  # no seed key -> 12345 (universe.py:272-277)
    uni.explode_grb(-50.0, 100.0)
    sd = uni.to_stan_data(-10.0, 40.0, dt=0.2, k=50)
    assert sd["N_detectors"] == 2 and sd["counts"].shape == (2, 249)
    rng = np.random.default_rng(1)
    S, k = 6, 50
    draws = [draw_prior(rng, k, 50.0) for _ in range(S)]
    fit = Fit(beta1=np.stack([d[1] for d in draws], 1) * 0.3, beta2=np.stack([d[2] for d in draws], 1) * 0.3,
              omega=np.stack([d[0] for d in draws], 2), scale=np.stack([d[3] for d in draws], 1),
              background=np.full((2, S), 500.0), amplitude=np.exp(rng.normal(0, 0.1, (1, S))) * 1.0,
              dt=np.full((1, S), uni.T0[1]))
    # give the intensity a realistic level: fold an overall factor of ~300 cps into detector 0 via its "amplitude"
    w, a, b = fit.folded()
    amp = np.full(S, 300.0)
    delays = np.linspace(-1.0, 1.0, 33)
    for d in range(2):
        nb = sd["N_time_bins"][d]
        ll = ctx.loglik_delay_grid(sd["time"][d, :nb], sd["exposure"][d, :nb], sd["counts"][d, :nb], w, a, b, amp,
                                   np.full(S, 500.0), delays)
        exp = orc.loglik_delay_grid(sd["counts"][d, :nb], sd["time"][d, :nb], sd["exposure"][d, :nb], w, a, b, amp,
                                    np.full(S, 500.0), delays)
        assert np.abs(ll - exp).max() < ATOL  # ~100-300 counts per 0.2 s bin: the fp64 epilogue is auto-selected
        np.testing.assert_array_equal(np.argmax(ll, axis=0), np.argmax(exp, axis=0))
    # expected_rate through the reference-named facade, detector 1 is shifted by dt and scaled by amplitude
    rate = fit.expected_rate(sd["time"][1, :249], 1)
    ref = orc.expected_rate_multiscale(sd["time"][1, :249], fit._omega1, fit._omega2, fit._beta1, fit._beta2, np.ones(S),
                                       fit._scale, fit._amplitude[0], fit._dt[0], S)
    np.testing.assert_allclose(rate, ref, rtol=1e-6)
    # sky grid over the two detectors (only the delay between them matters)
    from pyipn_b200.synth import healpix_ring_vectors

    ghat = healpix_ring_vectors(2)
    amp2 = np.stack([amp, amp * fit._amplitude[0]])
    sky = ctx.loglik_sky_grid(sd["N_time_bins"], sd["time"], sd["exposure"], sd["counts"], sd["sc_pos"], ghat, w, a, b, amp2,
                              np.full((2, S), 500.0))
    exp = orc.loglik_sky_grid(sd["counts"], sd["time"], sd["exposure"], sd["N_time_bins"], sd["sc_pos"], ghat, w, a, b, amp2,
                              np.full((2, S), 500.0))
    assert np.abs(sky - exp).max() < ATOL


def test_config4_sky_grid_full_size(ctx):
    from pyipn_b200.synth import sky_grid_workload

    w = sky_grid_workload(D=8, T=100_000, S=1, nside=64)
    assert w["ghat"].shape == (49152, 3)
    ll = ctx.loglik_sky_grid(w["nbins"], w["time"], w["exposure"], w["counts"], w["sc_pos"], w["ghat"], w["omega"], w["a"],
                             w["b"], w["amp"], w["bkg"])
    assert ll.shape == (49152, 1) and np.all(np.isfinite(ll))
    best = int(np.argmax(ll[:, 0]))
    # the best pixel is the one closest to the true direction (the delays of a solar-system-scale IPN are huge)
    assert np.dot(w["ghat"][best], w["g_true"]) == pytest.approx(np.max(w["ghat"] @ w["g_true"]), abs=1e-12)
    rng = np.random.default_rng(0)
    idx = np.unique(np.concatenate([[best], rng.choice(49152, 11, replace=False)]))
    from oracle import c_oracle as corc  # the plain-C twin of the numpy oracle (pinned to it in test_c_oracle.py): 100x faster

    corc.build()
    exp = corc.loglik_sky_grid(w["counts"], w["time"], w["exposure"], w["nbins"], w["sc_pos"], w["ghat"][idx], w["omega"], w["a"],
                               w["b"], w["amp"], w["bkg"])
    # 1e-3 nats absolute on the 8-detector sum for EVERY checked pixel: the random ones sit 1e5 - 1e6 nats below the maximum
    # (every far detector's model is seconds to minutes off there), which is where the fifth W digit plane and the
    # per-detector share of the tolerance (n_terms) are needed
    assert np.abs(ll[idx] - exp)[:, 0].max() < ATOL
    assert idx[int(np.argmax(exp[:, 0]))] == best
    # the device-resident entry point (ipn_loglik_sky_grid_dev) on the same inputs: bit-identical work, no host staging
    import torch

    dev = {k: torch.from_numpy(np.ascontiguousarray(w[k])).cuda() for k in ("time", "exposure", "counts", "sc_pos", "ghat", "omega", "a", "b", "amp", "bkg")}
    out = torch.empty((49152, 1), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    ctx.loglik_sky_grid_dev(w["nbins"], w["time"].shape[1], dev["time"].data_ptr(), dev["exposure"].data_ptr(), dev["counts"].data_ptr(),
                            dev["sc_pos"].data_ptr(), 49152, dev["ghat"].data_ptr(), 1, w["omega"].shape[1], dev["omega"].data_ptr(),
                            dev["a"].data_ptr(), dev["b"].data_ptr(), dev["amp"].data_ptr(), dev["bkg"].data_ptr(), out.data_ptr())
    ctx.synchronize()
    assert np.abs(out.cpu().numpy() - ll).max() < 1e-7


def test_config5_population_reduced(ctx):
    """GRB loop of config 5 with every stage on the GPU: intensity -> Poisson counts -> delay grid vs detector 0."""
    from pyipn_b200.params import fold_draws
    from pyipn_b200.synth import draw_prior

    rng = np.random.default_rng(5)
    T, G, k = 20_000, 512, 50
    mid = 1e-3 * (np.arange(T) + 0.5)
    expo = np.full(T, 1e-3)
    delays = np.linspace(-0.5, 0.5, G)
    for grb in range(3):
        om, b1, b2, sc = draw_prior(rng, k, mid[-1] - mid[0])
        w, a, b = fold_draws(om[0], om[1], b1, b2, 1.0, sc[:, None])
        true = np.concatenate([[0.0], rng.uniform(-0.3, 0.3, 2)])
        amp = 400.0 * np.exp(rng.normal(0, 0.2, 3))
        # per-detector expected source rate on the GPU (drop-in for _expeced_rate*): shift = true delay
        rates = ctx.rff_eval(mid, np.repeat(w, 3, 0), np.repeat(a, 3, 0), np.repeat(b, 3, 0), true, amp)
        counts = ctx.poisson_counts(expo, rates, np.full(3, 500.0), seed=100 + grb)
        assert counts.shape == (3, T) and abs(counts.mean() / ((rates.mean() + 500.0) * 1e-3) - 1.0) < 0.02
        for d in (1, 2):
            ll = ctx.loglik_delay_grid(mid, expo, counts[d], w, a, b, amp[d:d + 1], [500.0], delays)
            best = int(np.argmax(ll[:, 0]))
            sub = np.unique(np.clip([best - 1, best, best + 1, 0, G - 1], 0, G - 1))
            exp = orc.loglik_delay_grid(counts[d], mid, expo, w, a, b, amp[d:d + 1], [500.0], delays[sub])
            assert np.abs(ll[sub] - exp).max() < ATOL
            assert abs(delays[best] - true[d]) < 0.02  # a bright burst localises the delay to a few grid steps


def test_device_binning_is_bit_identical_to_numpy_histogram(ctx):
    """Next-row N3: ipn_bin_events == np.histogram(times, np.arange(tstart, tstop, dt)) incl. edge conventions."""
    from pyipn_b200 import LightCurve

    rng = np.random.default_rng(8)
    for tstart, tstop, dt in ((-10.0, 40.0, 0.2), (0.0, 100.0, 1e-3), (-3.3, 7.7, 0.037)):
        edges = np.arange(tstart, tstop, dt)
        times = np.concatenate([rng.uniform(tstart - 1.0, tstop + 1.0, 200_000), edges, edges[1:] - 1e-12, edges[:-1] + 1e-12,
                                [edges[-1], np.nextafter(edges[-1], np.inf), np.nan]])
        lc = LightCurve(times[:1000], times[1000:])
        rate, e, counts = lc.get_binned_light_curve(tstart, tstop, dt, device=True)
        ref_rate, ref_e, ref_counts = lc.get_binned_light_curve(tstart, tstop, dt)
        np.testing.assert_array_equal(counts, ref_counts)
        np.testing.assert_array_equal(e, ref_e)
        np.testing.assert_array_equal(rate, ref_rate)
    assert ctx.bin_events(np.zeros(0), 0.0, 1.0, 0.1)[2].sum() == 0
