"""GPU parity: delay-grid / sky-grid Poisson log-likelihood vs the float64 oracle composition.

Tolerances (BASELINE.json north_star): absolute 1e-3 nats on log-likelihood sums, argmax indices exact."""
import numpy as np
import pytest

from oracle import ipn_oracle as orc

pytestmark = pytest.mark.gpu
ATOL = 1e-3


def _fold_golden(g):
    w, a, b = zip(*[orc.fold_params(g["omega1"][:, s], g["omega2"][:, s], g["beta1"][:, s], g["beta2"][:, s], 1.0,
                                    g["scale"][0, s], g["scale"][1, s]) for s in range(g["omega1"].shape[1])])
    return np.array(w), np.array(a), np.array(b)


def test_delay_grid_reference_golden(ctx, loglik_golden):
    """Default path (tcgen05 sliced-integer contraction) against the fixture minted from the reference's rff.py."""
    g = loglik_golden
    w, a, b = _fold_golden(g)
    ll = ctx.loglik_delay_grid(g["time"], g["exposure"], g["counts"], w, a, b, g["amp"], g["bkg"], g["delays"])
    assert ll.shape == g["ll"].shape
    assert np.abs(ll - g["ll"]).max() < ATOL
    np.testing.assert_array_equal(np.argmax(ll, axis=0), np.argmax(g["ll"], axis=0))


@pytest.mark.parametrize("hifi", [1, 2, 3])
def test_delay_grid_epilogue_modes(ctx, loglik_golden, hifi):
    """All epilogues of the tcgen05 path (fp64 = 1, fp32 = 2, compensated fp32 = 3) meet the tolerance on the fixture
    (~27 counts/bin, 600 bins)."""
    from pyipn_b200._lib import OPT_HIFI

    g = loglik_golden
    w, a, b = _fold_golden(g)
    ctx.set_option(OPT_HIFI, hifi)
    try:
        ll = ctx.loglik_delay_grid(g["time"], g["exposure"], g["counts"], w, a, b, g["amp"], g["bkg"], g["delays"])
    finally:
        ctx.set_option(OPT_HIFI, 0)
    # both are bounded by the 2^-23 quantisation of the cos/sin table (|dx| ~ 3e-7 per bin): the fp64 epilogue only
    # removes the fp32 rounding of the per-bin terms, which matters for hundreds of counts per bin
    assert np.abs(ll - g["ll"]).max() < ATOL


@pytest.mark.parametrize("precise", [1, 0])
def test_delay_grid_fp32_pipe_path(ctx, loglik_golden, precise):
    """The FP32-pipe contraction (IPN_OPT_LOGLIK_IMPL = 1) is kept as the plain-CUDA-core baseline.  Its exponents carry
    the rounding of 200-400 fp32 additions (~1.4e-6 each), so its documented floor is ~5e-3 nats here, not 1e-3; the
    argmax still agrees."""
    from pyipn_b200._lib import OPT_LOGLIK_IMPL, OPT_PRECISE_W

    g = loglik_golden
    w, a, b = _fold_golden(g)
    ctx.set_option(OPT_LOGLIK_IMPL, 1)
    ctx.set_option(OPT_PRECISE_W, precise)
    try:
        ll = ctx.loglik_delay_grid(g["time"], g["exposure"], g["counts"], w, a, b, g["amp"], g["bkg"], g["delays"])
    finally:
        ctx.set_option(OPT_PRECISE_W, 1)
        ctx.set_option(OPT_LOGLIK_IMPL, 0)
    assert np.abs(ll - g["ll"]).max() < 5e-3
    np.testing.assert_array_equal(np.argmax(ll, axis=0), np.argmax(g["ll"], axis=0))


def test_facade_fit_delay_grid(ctx, loglik_golden):
    from pyipn_b200 import Fit

    g = loglik_golden
    fit = Fit(beta1=g["beta1"], beta2=g["beta2"], omega=np.stack([g["omega1"], g["omega2"]]), scale=g["scale"],
              background=np.stack([g["bkg"], g["bkg"]]), amplitude=g["amp"][None, :], dt=np.zeros((1, 3)))
    ll = fit.delay_grid_loglik(1, g["time"], g["exposure"], g["counts"], g["delays"])
    assert np.abs(ll - g["ll"]).max() < ATOL


@pytest.mark.parametrize("T,G,S,k", [(1, 1, 1, 1), (63, 5, 2, 3), (64, 128, 1, 50), (65, 129, 3, 50), (1000, 300, 5, 50),
                                     (20_000, 64, 2, 50)])
def test_delay_grid_vs_oracle_ragged_shapes(ctx, T, G, S, k):
    rng = np.random.default_rng(T * 7 + G + S)
    edges = np.cumsum(rng.uniform(0.002, 0.01, T + 1))
    time = 0.5 * (edges[1:] + edges[:-1])
    expo = np.diff(edges)
    omega = rng.normal(size=(S, 2 * k)) * np.where(np.arange(2 * k) < k, 0.05, 2.0)
    a = rng.normal(size=(S, 2 * k)) * 0.12
    b = rng.normal(size=(S, 2 * k)) * 0.12
    amp = np.exp(rng.normal(5.0, 0.3, S))
    bkg = np.exp(rng.normal(np.log(500.0), 0.1, S))
    f = np.exp(orc.log_intensity_folded(time - 0.05, omega[0], a[0], b[0]))
    counts = rng.poisson(expo * (amp[0] * f + bkg[0])).astype(np.int64)
    delays = np.sort(rng.uniform(-0.5, 0.5, G))
    ll = ctx.loglik_delay_grid(time, expo, counts, omega, a, b, amp, bkg, delays)
    exp = orc.loglik_delay_grid(counts, time, expo, omega, a, b, amp, bkg, delays)
    assert np.abs(ll - exp).max() < ATOL
    if G > 1:
        np.testing.assert_array_equal(np.argmax(ll, axis=0), np.argmax(exp, axis=0))


def test_delay_grid_edge_cases(ctx):
    rng = np.random.default_rng(2)
    T = 300
    time = np.arange(T) * 0.01
    expo = np.full(T, 0.01)
    omega = rng.normal(size=(2, 20))
    a = rng.normal(size=(2, 20)) * 0.2
    b = rng.normal(size=(2, 20)) * 0.2
    delays = np.linspace(-0.2, 0.2, 7)
    # all-zero counts; amplitude 0 in draw 1 (background only); tiny source in draw 0
    counts = np.zeros(T, dtype=np.int64)
    amp = np.array([1e-6, 0.0])
    bkg = np.array([500.0, 3.0])
    ll = ctx.loglik_delay_grid(time, expo, counts, omega, a, b, amp, bkg, delays)
    exp = orc.loglik_delay_grid(counts, time, expo, omega, a, b, amp, bkg, delays)
    assert np.abs(ll - exp).max() < 1e-6
    # bright source, ~500-5000 counts per bin (lgamma term matters; auto-selects the fp64 epilogue),
    # zero-exposure bins with zero counts
    amp = np.array([5e4, 4.8e4])
    bkg = np.array([500.0, 450.0])
    omega[1], a[1], b[1] = omega[0], a[0] * 1.01, b[0] * 0.99  # a posterior-like neighbour of the generating draw
    f = np.exp(orc.log_intensity_folded(time, omega[0], a[0], b[0]))
    counts = rng.poisson(expo * (amp[0] * f + bkg[0])).astype(np.int64)
    expo2 = expo.copy()
    expo2[10:20] = 0.0
    counts[10:20] = 0
    ll = ctx.loglik_delay_grid(time, expo2, counts, omega, a, b, amp, bkg, delays)
    exp = orc.loglik_delay_grid(counts, time, expo2, omega, a, b, amp, bkg, delays)
    assert np.abs(ll - exp).max() < ATOL
    # empty grids / empty bins
    assert ctx.loglik_delay_grid(time, expo, counts, omega, a, b, amp, bkg, np.zeros(0)).shape == (0, 2)
    ll0 = ctx.loglik_delay_grid(np.zeros(0), np.zeros(0), np.zeros(0, dtype=np.int64), omega, a, b, amp, bkg, delays)
    np.testing.assert_array_equal(ll0, 0.0)


def test_delay_grid_rejects_bad_input(ctx):
    from pyipn_b200 import IpnError

    t = np.arange(10.0)
    e = np.ones(10)
    n = np.ones(10, dtype=np.int64)
    w = np.ones((1, 4))
    with pytest.raises(IpnError):
        ctx.loglik_delay_grid(t, e, n, w, w, w, [1.0], [0.0], [0.0])  # bkg must be > 0
    with pytest.raises(IpnError):
        ctx.loglik_delay_grid(t, e, -n, w, w, w, [1.0], [1.0], [0.0])  # negative counts
    with pytest.raises(IpnError):
        ctx.loglik_delay_grid(t, e, n, w * np.nan, w, w, [1.0], [1.0], [0.0])


def test_config2_full_size_subsample_and_argmax(ctx):
    """BASELINE config 2 at full size (4096 delays x 1e5 bins, K=50, one draw): oracle on a 48-delay subsample
    (incl. the neighbourhood of the maximum), exact argmax, and the reported best-vs-second gap."""
    from pyipn_b200.synth import delay_grid_workload

    w = delay_grid_workload(T=100_000, G=4096, S=1)
    ll = ctx.loglik_delay_grid(w["time"], w["exposure"], w["counts"], w["omega"], w["a"], w["b"], w["amp"], w["bkg"], w["delays"])
    best = int(np.argmax(ll[:, 0]))
    assert abs(w["delays"][best] - w["true_delay"]) < 2.0 / 4096 + 2e-3
    idx = np.unique(np.concatenate([np.arange(0, 4096, 128), np.arange(best - 8, best + 8)]).clip(0, 4095))
    exp = orc.loglik_delay_grid(w["counts"], w["time"], w["exposure"], w["omega"], w["a"], w["b"], w["amp"], w["bkg"],
                                w["delays"][idx])
    err = np.abs(ll[idx, 0] - exp[:, 0]).max()
    assert err < ATOL, err
    assert idx[np.argmax(exp[:, 0])] == best
    srt = np.sort(ll[:, 0])
    assert srt[-1] - srt[-2] > 3 * err  # the argmax is decided by physics (gap ~3e-3 nats), not by rounding


def test_linearity_in_counts_full_size(ctx):
    """Size-independent property: the Poisson sum is linear in the counts vector apart from the lgamma term:
    ll(n1 + n2) + ll(0) - ll(n1) - ll(n2) = -sum[lgamma(n1+n2+1) - lgamma(n1+1) - lgamma(n2+1)] for every grid point."""
    from scipy.special import gammaln

    from pyipn_b200.synth import delay_grid_workload

    w = delay_grid_workload(T=100_000, G=256, S=2)
    rng = np.random.default_rng(0)
    n1 = w["counts"]
    n2 = rng.poisson(0.7, size=n1.shape).astype(np.int64)
    args = (w["omega"], w["a"], w["b"], w["amp"], w["bkg"], w["delays"])
    f = lambda n: ctx.loglik_delay_grid(w["time"], w["exposure"], n, *args)
    lhs = f(n1 + n2) + f(np.zeros_like(n1)) - f(n1) - f(n2)
    rhs = -(gammaln(n1 + n2 + 1.0) - gammaln(n1 + 1.0) - gammaln(n2 + 1.0)).sum()
    assert np.abs(lhs - rhs).max() < 4 * ATOL


def test_sky_grid_vs_oracle(ctx):
    from pyipn_b200.synth import sky_grid_workload

    w = sky_grid_workload(D=3, T=1500, S=2, nside=2, dt_bin=0.01)
    # ragged detectors: detector 2 has fewer valid bins (padding must be ignored)
    w["nbins"][2] = 1200
    ll = ctx.loglik_sky_grid(w["nbins"], w["time"], w["exposure"], w["counts"], w["sc_pos"], w["ghat"], w["omega"], w["a"],
                             w["b"], w["amp"], w["bkg"])
    exp = orc.loglik_sky_grid(w["counts"], w["time"], w["exposure"], w["nbins"], w["sc_pos"], w["ghat"], w["omega"], w["a"],
                              w["b"], w["amp"], w["bkg"])
    assert ll.shape == (48, 2)
    # nside = 2 pixels are ~30 degrees apart: most of them put the burst seconds away from where it is, ll drops by 1e4
    # nats and the residuals are tens of counts in every bin -- the tolerance is the same 1e-3 nats absolute there
    assert np.abs(ll - exp).max() < ATOL
    np.testing.assert_array_equal(np.argmax(ll, axis=0), np.argmax(exp, axis=0))
    # a finer look around the true direction: all of these are plausible pixels
    g0 = w["ghat"][int(np.argmax(exp[:, 0]))]
    rng = np.random.default_rng(4)
    close = g0 + rng.normal(size=(40, 3)) * 2e-3
    close /= np.linalg.norm(close, axis=1, keepdims=True)
    ll2 = ctx.loglik_sky_grid(w["nbins"], w["time"], w["exposure"], w["counts"], w["sc_pos"], close, w["omega"], w["a"],
                              w["b"], w["amp"], w["bkg"])
    exp2 = orc.loglik_sky_grid(w["counts"], w["time"], w["exposure"], w["nbins"], w["sc_pos"], close, w["omega"], w["a"],
                               w["b"], w["amp"], w["bkg"])
    assert np.abs(ll2 - exp2).max() < ATOL


@pytest.mark.parametrize("hook", [{"IPN_TC_GRID": "600"}, {"IPN_TC_GRID": "37"}, {"IPN_TC_EPI_DELAY": "20000"},
                                  {"IPN_TC_CHUNK_TILES": "16"}, {"IPN_TC_ONE_ISSUER": "1"}])
def test_tc_pipeline_schedule_independence(hooks_ctx, monkeypatch, hook):
    """The tcgen05 kernel's result must not depend on how its warp roles are paced or its items are cut: several CTA
    waves per SM, fewer CTAs than SMs, an artificially slow epilogue, tiny items, a single MMA issuer.  (A dead-lock in
    the barrier protocol shows up here as a launch failure with the time-out record in the message.)"""
    from pyipn_b200._lib import OPT_LOGLIK_IMPL
    from pyipn_b200.synth import delay_grid_workload

    ctx = hooks_ctx
    w = delay_grid_workload(T=30_000, G=1024, S=6)
    args = (w["time"], w["exposure"], w["counts"], w["omega"], w["a"], w["b"], w["amp"], w["bkg"], w["delays"])
    ctx.set_option(OPT_LOGLIK_IMPL, 2)
    try:
        base = ctx.loglik_delay_grid(*args)
        for k, v in hook.items():
            monkeypatch.setenv(k, v)
        got = ctx.loglik_delay_grid(*args)
    finally:
        ctx.set_option(OPT_LOGLIK_IMPL, 0)
    # same arithmetic per (delay, bin); only the order of the float64 partial sums changes
    assert np.abs(got - base).max() < 1e-7
    exp = orc.loglik_delay_grid(w["counts"], w["time"], w["exposure"], w["omega"][:1], w["a"][:1], w["b"][:1], w["amp"][:1],
                                w["bkg"][:1], w["delays"][::64])
    assert np.abs(got[::64, :1] - exp).max() < ATOL


@pytest.mark.parametrize("hook", [{"IPN_TC_SKIP_MATH": "1"}, {"IPN_TC_SKIP_MATH": "2"}, {"IPN_TC_SKIP_MMA": "1"}])
def test_tc_pipeline_survives_unbalanced_roles(hooks_ctx, monkeypatch, hook):
    """Experiment switches that make one role much faster than the others (values are meaningless then): the hand-shakes
    must still complete.  This is the regime in which a skipped barrier phase dead-locked the two-issuer kernel."""
    from pyipn_b200._lib import OPT_LOGLIK_IMPL
    from pyipn_b200.synth import delay_grid_workload

    ctx = hooks_ctx
    w = delay_grid_workload(T=100_000, G=2048, S=8)
    for k, v in hook.items():
        monkeypatch.setenv(k, v)
    ctx.set_option(OPT_LOGLIK_IMPL, 2)
    try:
        ll = ctx.loglik_delay_grid(w["time"], w["exposure"], w["counts"], w["omega"], w["a"], w["b"], w["amp"], w["bkg"], w["delays"])
    finally:
        ctx.set_option(OPT_LOGLIK_IMPL, 0)
    assert ll.shape == (2048, 8)


def test_shipped_library_ignores_the_hook_switches(ctx, monkeypatch):
    """The shipped library has no hooks: IPN_TC_SKIP_MATH in a user's environment must not change a single bit."""
    from pyipn_b200.synth import delay_grid_workload

    w = delay_grid_workload(T=6000, G=256, S=3)
    args = (w["time"], w["exposure"], w["counts"], w["omega"], w["a"], w["b"], w["amp"], w["bkg"], w["delays"])
    base = ctx.loglik_delay_grid(*args)
    monkeypatch.setenv("IPN_TC_SKIP_MATH", "1")
    monkeypatch.setenv("IPN_TC_SKIP_MMA", "1")
    got = ctx.loglik_delay_grid(*args)
    exp = orc.loglik_delay_grid(w["counts"], w["time"], w["exposure"], w["omega"], w["a"], w["b"], w["amp"], w["bkg"], w["delays"][::16])
    assert np.abs(got[::16] - exp).max() < ATOL and np.abs(got - base).max() < 1e-7


@pytest.mark.parametrize("boost,comp_guaranteed", [(3_000.0, True), (20_000.0, False)])
def test_compensated_epilogue_on_bright_data(ctx, boost, comp_guaranteed):
    """1e5 bins with ~10 (sum n^2 = 2.3e7: auto picks the compensated fp32 epilogue) or ~40 counts each (sum n^2 = 6e8: auto
    picks fp64).  The plain fp32 epilogue rounds its per-bin terms at 6e-8 n and is not guaranteed here; the compensated
    one is up to sum n^2 = 4e7; auto mode must meet the tolerance in both."""
    from oracle import c_oracle as corc
    from pyipn_b200._lib import OPT_HIFI, OPT_LOGLIK_IMPL
    from pyipn_b200.synth import delay_grid_workload

    w = delay_grid_workload(T=100_000, G=512, S=2, boost=boost, bkg=5_000.0)
    sumsq = float((w["counts"].astype(np.float64) ** 2).sum())
    assert (1e7 < sumsq < 4e7) if comp_guaranteed else (2e8 < sumsq < 2e9)
    args = (w["time"], w["exposure"], w["counts"], w["omega"], w["a"], w["b"], w["amp"], w["bkg"], w["delays"])
    idx = np.arange(0, 512, 8)
    exp = corc.loglik_delay_grid(w["counts"], w["time"], w["exposure"], w["omega"], w["a"], w["b"], w["amp"], w["bkg"],
                                 w["delays"][idx])
    errs = {}
    ctx.set_option(OPT_LOGLIK_IMPL, 2)
    try:
        for mode in (0, 3, 1, 2):
            ctx.set_option(OPT_HIFI, mode)
            ll = ctx.loglik_delay_grid(*args)
            # every checked (delay, draw) at 1e-3 nats absolute, however wrong the model is there -- up to a deficit of 1e6
            # nats.  Beyond that (only the jittered second draw of the ~40 counts / bin data set: 4e6 nats below the best
            # point, residuals of hundreds of counts in every bin) the 2^-39 rounding of W, which adds up coherently over the
            # bins, is the floor: 2e-9 of the deficit (measured 1.3e-9)
            deficit = exp.max() - exp
            errs[mode] = float((np.abs(ll[idx] - exp) / np.where(deficit < 1.0e6, 1.0, 2.0e-6 * deficit)).max())
            np.testing.assert_array_equal(np.argmax(ll[idx], axis=0), np.argmax(exp, axis=0))
    finally:
        ctx.set_option(OPT_HIFI, 0)
        ctx.set_option(OPT_LOGLIK_IMPL, 0)
    print(f"sum n^2 = {sumsq:.3g}: max |dll| by epilogue mode (0 auto, 3 compensated, 1 fp64, 2 plain fp32):", errs)
    assert errs[0] < ATOL and errs[1] < ATOL
    if comp_guaranteed:
        assert errs[3] < ATOL


@pytest.mark.parametrize("impl", [2, 1])
def test_delay_grid_random_small_shapes_property(ctx, impl):
    """Hypothesis over small ragged shapes (bins not a multiple of 32, delays not a multiple of 128, J from 1 to 120, zero
    exposure, zero amplitude, empty light curves): both implementations against the C oracle."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    from oracle import c_oracle as corc
    from pyipn_b200._lib import OPT_LOGLIK_IMPL

    @settings(max_examples=25, deadline=None, derandomize=True, database=None)
    # the tcgen05 path holds 2 J + 8 rows in its K = 224 digit planes: J <= 108 (auto mode takes the FP32-pipe kernel beyond)
    @given(T=st.integers(0, 200), G=st.integers(1, 300), S=st.integers(1, 4), J=st.integers(1, 108 if impl == 2 else 120),
           seed=st.integers(0, 2**31 - 1))
    def check(T, G, S, J, seed):
        rng = np.random.default_rng(seed)
        time = np.sort(rng.uniform(0, 20, T))
        expo = rng.uniform(0.001, 0.05, T)
        if T > 3:
            expo[2] = 0.0
        omega = rng.normal(size=(S, J)) * np.where(np.arange(J) % 2 == 0, 0.1, 3.0)
        a = rng.normal(size=(S, J)) * 0.1
        b = rng.normal(size=(S, J)) * 0.1
        amp = np.exp(rng.normal(5.0, 0.5, S))
        if S > 1:
            amp[1] = 0.0
        bkg = np.exp(rng.normal(np.log(300.0), 0.3, S))
        counts = rng.poisson(expo * 600.0).astype(np.int64)
        delays = np.sort(rng.uniform(-1, 1, G))
        got = ctx.loglik_delay_grid(time, expo, counts, omega, a, b, amp, bkg, delays)
        exp = corc.loglik_delay_grid(counts, time, expo, omega, a, b, amp, bkg, delays)
        assert got.shape == (G, S)
        assert np.abs(got - exp).max() < ATOL

    ctx.set_option(OPT_LOGLIK_IMPL, impl)
    try:
        check()
        if impl == 2:  # forced tcgen05 with too many frequencies is an explicit error, not a silent switch
            from pyipn_b200 import IpnError

            z = np.zeros((1, 109))
            with pytest.raises(IpnError):
                ctx.loglik_delay_grid(np.arange(4.0), np.ones(4), np.zeros(4, dtype=np.int64), z, z, z, [1.0], [1.0], [0.0])
    finally:
        ctx.set_option(OPT_LOGLIK_IMPL, 0)
    z = np.zeros((1, 109))
    assert ctx.loglik_delay_grid(np.arange(4.0), np.ones(4), np.zeros(4, dtype=np.int64), z, z, z, [1.0], [1.0], [0.0]).shape == (1, 1)


@pytest.mark.parametrize("impl", [2, 1])
def test_draw_batches_are_stitched_correctly(ctx, impl):
    """The library works through the draws in internal batches (bounded operand-image scratch).  With IPN_OPT_DRAW_BATCH = 2
    and S = 5 the call crosses two batch boundaries and ends on a ragged batch: every draw must match the oracle
    (functions.stan:23-35 composed with fit.py:671-691) and the un-batched call bit for bit (each draw's arithmetic does not
    depend on its batch)."""
    from pyipn_b200._lib import OPT_DRAW_BATCH, OPT_LOGLIK_IMPL
    from pyipn_b200.synth import delay_grid_workload

    w = delay_grid_workload(T=7001, G=200, S=5)
    args = (w["time"], w["exposure"], w["counts"], w["omega"], w["a"], w["b"], w["amp"], w["bkg"], w["delays"])
    ctx.set_option(OPT_LOGLIK_IMPL, impl)
    try:
        whole = ctx.loglik_delay_grid(*args)
        assert ctx.last_draw_batch == 5
        ctx.set_option(OPT_DRAW_BATCH, 2)
        got = ctx.loglik_delay_grid(*args)
        assert ctx.last_draw_batch == 2
    finally:
        ctx.set_option(OPT_DRAW_BATCH, 0)
        ctx.set_option(OPT_LOGLIK_IMPL, 0)
    exp = orc.loglik_delay_grid(w["counts"], w["time"], w["exposure"], w["omega"], w["a"], w["b"], w["amp"], w["bkg"], w["delays"])
    assert np.abs(got - exp).max() < ATOL
    np.testing.assert_array_equal(np.argmax(got, axis=0), np.argmax(exp, axis=0))
    if impl == 2:  # the sliced-integer contraction is exact and the per-thread sums are deterministic per (draw, delay)
        np.testing.assert_allclose(got, whole, rtol=0, atol=2e-9)


def test_many_draws_full_length_vs_oracle(ctx):
    """Bench shape in the draw and bin dimensions (S = 200 > one internal batch of ~85 draws, T = 1e5, fewer delays): a
    subsample of (delay, draw) pairs from EVERY internal batch, including their first and last draws, against the C oracle
    at 1e-3 nats, per-draw arg-max on a window around the device's maximum."""
    from oracle import c_oracle as corc
    from pyipn_b200.synth import delay_grid_workload

    corc.build()
    S, G = 200, 512
    w = delay_grid_workload(T=100_000, G=G, S=S)
    ll = ctx.loglik_delay_grid(w["time"], w["exposure"], w["counts"], w["omega"], w["a"], w["b"], w["amp"], w["bkg"], w["delays"])
    batch = ctx.last_draw_batch
    assert 1 <= batch < S, "the call must span several internal batches for this test to mean anything"
    draws = sorted({0, S - 1, *range(0, S, batch), *(min(S, b + batch) - 1 for b in range(0, S, batch)), 17, 101})
    best = int(np.argmax(ll[:, 0]))
    gsel = sorted({0, best, min(G - 1, best + 1), G - 1})
    exp = corc.loglik_delay_grid(w["counts"], w["time"], w["exposure"], w["omega"][draws], w["a"][draws], w["b"][draws],
                                 w["amp"][draws], w["bkg"][draws], w["delays"][gsel])
    assert np.abs(ll[np.ix_(gsel, draws)] - exp).max() < ATOL
    for s in (draws[1], draws[-2]):
        b = int(np.argmax(ll[:, s]))
        win = np.arange(max(0, b - 2), min(G, b + 3))
        e = corc.loglik_delay_grid(w["counts"], w["time"], w["exposure"], w["omega"][[s]], w["a"][[s]], w["b"][[s]],
                                   w["amp"][[s]], w["bkg"][[s]], w["delays"][win])[:, 0]
        assert win[int(np.argmax(e))] == b


@pytest.mark.parametrize("T,G,S", [(1_000_003, 130, 2), (3001, 70_001, 1)])
def test_large_shapes_vs_oracle(ctx, T, G, S):
    """A million bins (31 251 tiles, ragged last one, 245 items per delay tile) and seventy thousand delays (547 delay tiles, the
    last one with a single row): a subsample of (delay, draw) pairs against the C oracle, arg-max on a window."""
    from oracle import c_oracle as corc
    from pyipn_b200.synth import delay_grid_workload

    corc.build()
    w = delay_grid_workload(T=T, G=G, S=S, dt_bin=1e-4 if T > 100_000 else 1e-2, bkg=5000.0 if T > 100_000 else 50.0)
    ll = ctx.loglik_delay_grid(w["time"], w["exposure"], w["counts"], w["omega"], w["a"], w["b"], w["amp"], w["bkg"], w["delays"])
    assert ll.shape == (G, S) and np.all(np.isfinite(ll))
    best = int(np.argmax(ll[:, 0]))
    idx = np.unique(np.clip([0, best - 1, best, best + 1, G // 3, G - 2, G - 1], 0, G - 1))
    exp = corc.loglik_delay_grid(w["counts"], w["time"], w["exposure"], w["omega"], w["a"], w["b"], w["amp"], w["bkg"], w["delays"][idx])
    assert np.abs(ll[idx] - exp).max() < ATOL
    # the device's arg-max is the oracle's, unless the grid is so fine (70 001 delays: 1.4e-5 s apart) that the oracle's best
    # points are closer together than the tolerance -- then it must be one of those
    e0 = exp[:, 0]
    assert idx[int(np.argmax(e0))] == best or e0[list(idx).index(best)] > e0.max() - ATOL
