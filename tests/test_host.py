"""CPU: host-side logic of the facade (parameter folding, binning containers, Universe data contract, generators)."""
import numpy as np
import pytest

from oracle import ipn_oracle as orc


def test_fold_draws_matches_oracle_fold(rff_golden):
    from pyipn_b200 import fold_draws

    g = rff_golden
    w, a, b = fold_draws(g["B_omega1"], g["B_omega2"], g["B_beta1"], g["B_beta2"], g["B_bw"], g["B_scale"])
    assert w.shape == (7, 100) and w.flags.c_contiguous
    for s in range(7):
        ow, oa, ob = orc.fold_params(g["B_omega1"][:, s], g["B_omega2"][:, s], g["B_beta1"][:, s], g["B_beta2"][:, s],
                                     g["B_bw"][s], g["B_scale"][0, s], g["B_scale"][1, s])
        np.testing.assert_array_equal(w[s], ow)
        np.testing.assert_array_equal(a[s], oa)
        np.testing.assert_array_equal(b[s], ob)
    # single scale == both scales equal; 1-D inputs are one draw
    w1, a1, b1 = fold_draws(g["B_omega1"][:, 0], g["B_omega2"][:, 0], g["B_beta1"][:, 0], g["B_beta2"][:, 0], 1.0, 0.3)
    assert w1.shape == (1, 100)
    np.testing.assert_allclose(a1[0, :50], a1[0, 50:])
    with pytest.raises(ValueError):
        fold_draws(np.ones((3, 2)), np.ones((3, 2)), np.ones((4, 2)), np.ones((3, 2)), 1.0, 1.0)


def test_lightcurve_binning_matches_reference_semantics():
    from pyipn_b200 import BinnedLightCurve, LightCurve

    rng = np.random.default_rng(0)
    src, bkg = np.sort(rng.uniform(0, 10, 500)), np.sort(rng.uniform(-5, 20, 2000))
    lc = LightCurve(src, bkg)
    rate, edges, counts = lc.get_binned_light_curve(-2.0, 12.0, 0.25)
    r2, e2, c2 = orc.bin_events(np.concatenate([src, bkg]), -2.0, 12.0, 0.25)
    np.testing.assert_array_equal(counts, c2)
    np.testing.assert_array_equal(edges, e2)
    blc = BinnedLightCurve.from_lightcurve(lc, -2.0, 12.0, 0.25)
    assert blc.n_bins == len(counts) and blc.res_ms == 250.0
    np.testing.assert_allclose(blc.exposure, 0.25)
    assert blc.time2idx(0.0) == np.searchsorted(edges, 0.0)
    # get_max_sn: python loop of the reference (lightcurve.py:146-173) restated literally
    len_int = int(1000.0 // blc.res_ms)
    cs = np.cumsum(blc.get_src_counts(500.0 * 0 + 80.0))
    fmax, i1, i2 = 0.0, 0, cs.size
    for i in range(cs.size - len_int):
        dif = cs[i + len_int] - cs[i]
        if dif > fmax:
            fmax, i1, i2 = dif, i, i + len_int
    assert blc.get_max_sn(1000.0, bkg_rate=80.0) == (i1, i2)
    with pytest.raises(AssertionError):
        BinnedLightCurve(np.zeros(3), np.zeros(3), 0, 1, 0.1)


def test_generators_are_seeded_and_thin_correctly():
    from pyipn_b200.possion_gen import background_poisson_generator, norris, source_poisson_generator

    a = background_poisson_generator(0.0, 20.0, 0.0, 500.0, seed=5)
    b = background_poisson_generator(0.0, 20.0, 0.0, 500.0, seed=5)
    np.testing.assert_array_equal(a, b)
    assert a[0] == 0.0 and np.all(np.diff(a) > 0)
    assert abs(len(a) - 500 * 20) < 5 * np.sqrt(500 * 20)
    # literal restatement of the reference loop (possion_gen.py:232-240) consumes the same RandomState stream
    rs = np.random.RandomState(5)
    t, ref = 0.0, [0.0]
    while t < 20.0:
        t = t - (1.0 / 500.0) * np.log(rs.rand())
        if rs.rand() <= 1.0:
            ref.append(t)
    np.testing.assert_allclose(a, np.array(ref), rtol=0, atol=1e-9)
    s = source_poisson_generator(-5.0, 30.0, 300.0, 0.0, 0.5, 4.0, seed=7)
    assert s[0] == -5.0 and np.all(s[1:] > 0.0)  # no source photons before the pulse starts
    x = np.linspace(-5, 30, 7001)
    expected = np.trapezoid(norris(x, 300.0, 0.0, 0.5, 4.0), x)
    assert abs((len(s) - 1) - expected) < 6 * np.sqrt(expected)
    assert len(source_poisson_generator(0.0, 5.0, 300.0, 100.0, 0.5, 4.0, seed=1)) == 0  # pulse after the window


def test_universe_from_dict_and_stan_data_contract():
    from pyipn_b200 import Universe

    d = dict(
        seed=42,
        grb=dict(ra=180.0, dec=30.0, distance=500.0, K=500.0, t_rise=0.5, t_decay=4.1),
        detectors=dict(
            det3=dict(ra=300.0, dec=-45.0, altitude=153000.0, time="2010-01-01T00:00:00", pointing=dict(ra=180.0, dec=30.0),
                      effective_area=1.0),
            det2=dict(ra=60.0, dec=10.0, altitude=500.0, time="2010-01-01T00:00:00", pointing=dict(ra=180.0, dec=30.0),
                      effective_area=1.0),
        ),
    )
    u = Universe.from_dict(d)
    u.explode_grb(-10.0, 30.0)
    assert u.T0[0] == 0.0 and 0.30 < u.T0[1] < 0.40  # notebook: 0.34328823 with astropy frames, 0.342937 plain trig
    assert list(u.detectors.keys())[0] == "det2"  # re-ordered by arrival time
    sd = u.to_stan_data(-5.0, 20.0, dt=0.2, k=50, n_cores=4)
    nb = len(np.arange(-5.0, 20.0, 0.2)) - 1
    assert sd["N_detectors"] == 2 and sd["N_time_bins"] == [nb, nb] and sd["max_N_time_bins"] == nb
    assert sd["counts"].dtype.kind == "i" and sd["counts"].shape == (2, nb)
    np.testing.assert_allclose(sd["exposure"], 0.2)
    np.testing.assert_allclose(np.linalg.norm(sd["sc_pos"], axis=1), [6371.0 + 500.0, 6371.0 + 153000.0])
    assert sd["grainsize"] == [int(np.round(nb / 4) * 0.5)] * 2 and sd["grainsize_meta"] == 1
    # the delay the Stan model would infer from sc_pos has the same magnitude as T0 (different c constants: 1e-8 rel)
    g = u.grb.norm_vec
    dt = orc.time_delay(g, sd["sc_pos"][0] - sd["sc_pos"][1])
    assert abs(abs(dt) - u.T0[1]) < 1e-6
    # same seed -> same light curves
    u2 = Universe.from_dict(d)
    u2.explode_grb(-10.0, 30.0)
    np.testing.assert_array_equal(u2.to_stan_data(-5.0, 20.0)["counts"], sd["counts"])


def test_universe_from_yaml_template(tmp_path):
    """The schema of the reference's examples/template_3.yaml (BASELINE config 1): no seed key -> 12345 (universe.py:272-277),
    detectors registered in file order and re-ordered by arrival time when the GRB goes off."""
    from pyipn_b200 import Universe

    text = """
grb:
  ra: 180.
  dec: 30.
  distance: 500.
  K: 500. # intensity
  t_rise: .5
  t_decay: 4.1
detectors:
  det3:
     ra: 300.
     dec: -45.
     altitude: 153000.
     time: '2010-01-01T00:00:00'
     pointing:
       ra: 180.
       dec: 30.
     effective_area: 1.
  det2:
    ra: 60.
    dec: 10.
    altitude: 500  # km
    time: '2010-01-01T00:00:00'
    pointing:
      ra: 180.
      dec: 30.
    effective_area: 1.
"""
    f = tmp_path / "template.yaml"
    f.write_text(text)
    u = Universe.from_yaml(str(f))
    assert u._seed == 12345 and list(u.detectors.keys()) == ["det3", "det2"]
    u.explode_grb(-5.0, 10.0)
    assert list(u.detectors.keys()) == ["det2", "det3"] and u.T0[0] == 0.0 and 0.30 < u.T0[1] < 0.40
    sd = u.to_stan_data(-2.0, 8.0, dt=0.2, k=50)
    assert sd["N_detectors"] == 2 and sd["counts"].shape == (2, len(np.arange(-2.0, 8.0, 0.2)) - 1) and sd["k"] == 50
    assert sd["counts"].sum() > 2 * 500 * 9  # 500 cps of background per detector plus the burst
    # pulse lists (several pulses per burst) are accepted by the same schema (universe.py:283-298)
    import yaml

    d = yaml.safe_load(text)
    d["grb"].update(K=[300.0, 150.0], t_rise=[0.3, 0.6], t_decay=[2.0, 1.5], t_start=[0.0, 2.5])
    u2 = Universe.from_dict(d)
    u2.explode_grb(-5.0, 10.0)
    assert len(u2.light_curves["det2"].source_arrival_times) > 100


def test_fit_argument_routing_without_gpu(monkeypatch):
    """Fit.expected_rate picks dt / amplitude exactly like pyipn/fit.py:249-267 (checked by intercepting the ABI call)."""
    import pyipn_b200.fit as fitmod

    S, k = 4, 3
    rng = np.random.default_rng(0)
    fit = fitmod.Fit(beta1=rng.normal(size=(k, S)), beta2=rng.normal(size=(k, S)), omega=rng.normal(size=(2, k, S)),
                     scale=np.abs(rng.normal(size=(2, S))), background=np.full((3, S), 500.0),
                     amplitude=np.arange(2 * S, dtype=float).reshape(2, S) + 1, dt=np.arange(2 * S, dtype=float).reshape(2, S))
    seen = {}

    class Fake:
        def rff_eval(self, time, omega, a, b, shift, amp):
            seen.update(shift=np.array(shift), amp=np.array(amp), omega=omega)
            return np.zeros((len(shift), len(time)))

    monkeypatch.setattr(fitmod, "default_context", lambda: Fake())
    fit.expected_rate(np.arange(5.0), 0)
    np.testing.assert_array_equal(seen["shift"], 0.0)
    np.testing.assert_array_equal(seen["amp"], 1.0)
    fit.expected_rate(np.arange(5.0), 2)
    np.testing.assert_array_equal(seen["shift"], [4.0, 5.0, 6.0, 7.0])
    np.testing.assert_array_equal(seen["amp"], [5.0, 6.0, 7.0, 8.0])
    assert seen["omega"].shape == (S, 2 * k)
    with pytest.raises(AssertionError):
        fit.expected_rate(np.arange(5.0), 3)


def test_universe_annuli_triangulation_and_accessors():
    """calculate_annulus / localize_GRB (universe.py:364-386, 670-697) and the accessors of GRB / Detector / Fit: the true
    direction lies on every annulus, and four detectors in general position recover it by least squares."""
    from pyipn_b200 import Detector, Fit, GRB, Universe

    grb = GRB(ra=80.0, dec=-30.0, distance=500.0, K=500.0, t_rise=0.5, t_decay=4.1)
    u = Universe(grb, seed=7)
    for name, ra, dec, alt in (("a", 10.0, 5.0, 500.0), ("b", 130.0, 40.0, 153000.0), ("c", 250.0, -50.0, 90000.0),
                               ("d", 300.0, 70.0, 40000.0)):
        u.register_detector(Detector(ra, dec, alt, 80.0, -30.0, 1.0, name))
    u._compute_time_differences()
    g = grb.norm_vec
    names = list(u.detectors.keys())
    for d1 in names:
        for d2 in names:
            if d1 == d2:
                continue
            n, radec, theta = u.calculate_annulus(d1, d2)
            assert abs(np.dot(n, g) - np.cos(theta)) < 1e-12
            np.testing.assert_allclose(n, [np.cos(np.radians(radec[1])) * np.cos(np.radians(radec[0])),
                                           np.cos(np.radians(radec[1])) * np.sin(np.radians(radec[0])),
                                           np.sin(np.radians(radec[1]))], atol=1e-12)
    sol = u.localize_GRB()
    np.testing.assert_allclose(sol / np.linalg.norm(sol), g, atol=1e-9)
    # accessors
    det = u.detectors["b"]
    assert det.effective_area == 1.0 and np.allclose(det.pointing, g) and np.allclose(det.location, det.xyz_km)
    assert abs(det.angular_separation(grb)) < 1e-7
    assert 5.0e16 < det.light_travel_time(grb) < 5.2e16  # 500 Mpc in light-seconds
    assert u.grb_radius == 500.0 and u.table["name"] == names and len(u.table["dt"]) == 4
    assert grb.table["K"] == 500.0 and np.allclose(grb.location / np.linalg.norm(grb.location), g)
    S, k = 3, 2
    rng = np.random.default_rng(1)
    arrays = dict(beta1=rng.normal(size=(k, S)), beta2=rng.normal(size=(k, S)), omega=rng.normal(size=(2, k, S)),
                  scale=np.ones((2, S)), background=np.full((2, S), 500.0), amplitude=np.ones((1, S)), dt=np.zeros((1, S)),
                  grb_theta=np.zeros(S), grb_phi=np.ones(S))
    fit = Fit(**arrays)
    for key in ("beta1", "beta2", "scale", "background", "amplitude", "dt", "grb_theta", "grb_phi"):
        np.testing.assert_array_equal(getattr(fit, key), arrays[key])
    np.testing.assert_array_equal(fit.omega1, arrays["omega"][0])
    np.testing.assert_array_equal(fit.omega2, arrays["omega"][1])


def test_fit_callers_of_the_path_without_gpu(monkeypatch):
    """The numeric halves of plot_light_curve_fit / plot_light_curve_ppcs / _compute_ppcs (fit.py:447-579): which light
    curve, bins, background row and draws reach the two device calls (intercepted; the oracle stands in for the device)."""
    import pyipn_b200.fit as fitmod
    from pyipn_b200 import Universe

    d = dict(seed=3, grb=dict(ra=180.0, dec=30.0, distance=500.0, K=500.0, t_rise=0.5, t_decay=4.1),
             detectors=dict(det3=dict(ra=300.0, dec=-45.0, altitude=153000.0, pointing=dict(ra=180.0, dec=30.0), effective_area=1.0),
                            det2=dict(ra=60.0, dec=10.0, altitude=500.0, pointing=dict(ra=180.0, dec=30.0), effective_area=1.0)))
    u = Universe.from_dict(d)
    u.explode_grb(-5.0, 15.0)
    S, k = 40, 4
    rng = np.random.default_rng(5)
    fit = fitmod.Fit(beta1=0.3 * rng.normal(size=(k, S)), beta2=0.3 * rng.normal(size=(k, S)), omega=rng.normal(size=(2, k, S)),
                     scale=np.ones((2, S)), background=np.stack([np.full(S, 500.0), np.full(S, 450.0)]),
                     amplitude=np.full((1, S), 0.8), dt=np.full((1, S), 0.34), universe=u)

    class Fake:
        def rff_eval(self, time, omega, a, b, shift, amp):
            return np.stack([amp[s] * np.exp(orc.log_intensity_folded(time - shift[s], omega[s], a[s], b[s])) for s in range(len(amp))])

        def poisson_counts(self, exposure, rates, bkg, seed):
            self.seen = dict(exposure=np.array(exposure), bkg=np.array(bkg), seed=seed)
            return np.random.default_rng(seed).poisson((rates + bkg[:, None]) * exposure).astype(np.int64)

    fake = Fake()
    monkeypatch.setattr(fitmod, "default_context", lambda: fake)
    mid, rate, model = fit.light_curve_fit(1, -2.0, 8.0, dt=0.2, thin=4)
    lc = u.light_curves[list(u.light_curves.keys())[1]]
    r, edges, counts = lc.get_binned_light_curve(-2.0, 8.0, 0.2)
    np.testing.assert_array_equal(rate, r)
    np.testing.assert_allclose(mid, 0.5 * (edges[:-1] + edges[1:]))
    assert model.shape == (S // 4, mid.size)
    np.testing.assert_allclose(model[1], fit.expected_rate(mid, 1)[4] + 450.0)
    ppcs = fit._compute_ppcs(1, -2.0, 8.0, 0.2, seed=11)
    assert ppcs.shape == (S, mid.size) and ppcs.dtype == np.float64
    np.testing.assert_array_equal(fake.seen["bkg"], 450.0)  # detector 1's background row
    np.testing.assert_allclose(fake.seen["exposure"], np.diff(edges))
    e2, r2, bands = fit.light_curve_ppc_bands(1, -2.0, 8.0, dt=0.2, levels=(95, 68), seed=11)
    np.testing.assert_array_equal(e2, edges)
    (lo95, hi95), (lo68, hi68) = bands
    assert (lo95 <= lo68).all() and (lo68 <= hi68).all() and (hi68 <= hi95).all()
    np.testing.assert_allclose(lo68, np.percentile(ppcs / np.diff(edges), 16.0, axis=0))
    with pytest.raises(AssertionError):
        fitmod.Fit(beta1=np.zeros((k, S)), beta2=np.zeros((k, S)), omega=np.zeros((2, k, S)), scale=np.ones(S),
                   background=np.full(S, 500.0)).light_curve_fit(0, 0.0, 1.0)  # no universe attached
