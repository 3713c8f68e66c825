"""GPU parity: ipn_rff_eval (through the C ABI and the reference-named Python facade) vs the reference fixtures / oracle.

Tolerance (BASELINE.json north_star): relative 1e-6 on lambda."""
import numpy as np
import pytest

from oracle import ipn_oracle as orc

pytestmark = pytest.mark.gpu
RTOL = 1e-6


def test_rff_golden_single_and_multiscale(ctx, rff_golden):
    from pyipn_b200 import RFF, RFF_multiscale

    g = rff_golden
    got = RFF(g["A_time"], g["A_omega1"], g["A_omega2"], g["A_beta1"], g["A_beta2"], float(g["A_bw"]), g["A_scale"][1])
    np.testing.assert_allclose(got, g["A_rff"], rtol=RTOL)
    got = RFF_multiscale(g["A_time"], g["A_omega1"], g["A_omega2"], g["A_beta1"], g["A_beta2"], float(g["A_bw"]),
                         g["A_scale"][0], g["A_scale"][1])
    np.testing.assert_allclose(got, g["A_rff_ms"], rtol=RTOL)
    assert got.dtype == np.float64 and got.shape == g["A_time"].shape


def test_expected_rate_golden_strided_draws(ctx, rff_golden):
    from pyipn_b200 import _expeced_rate, _expeced_rate_multiscale

    g = rff_golden
    args = (g["B_time"], g["B_omega1"], g["B_omega2"], g["B_beta1"], g["B_beta2"], g["B_bw"])
    np.testing.assert_allclose(_expeced_rate(*args, g["B_scale"][1], g["B_amp"], g["B_dt"], 7), g["B_rate_single"], rtol=RTOL)
    np.testing.assert_allclose(_expeced_rate_multiscale(*args, g["B_scale"], g["B_amp"], g["B_dt"], 7), g["B_rate_multi"],
                               rtol=RTOL)


def test_rff_edge_cases(ctx, rff_golden):
    from pyipn_b200 import RFF, RFF_multiscale

    g = rff_golden
    got = RFF_multiscale(g["C_time"], g["C_omega1"], g["C_omega2"], g["C_beta1"], g["C_beta2"], 1.0, 0.4, 1.3)
    np.testing.assert_allclose(got, g["C_rff_ms"], rtol=RTOL)  # one bin, k = 1, |phase| ~ 9e3 rad
    np.testing.assert_array_equal(RFF(g["A_time"], np.zeros(3), np.zeros(3), np.zeros(3), np.zeros(3), 1.0, 1.0), 1.0)
    assert RFF(np.zeros(0), np.ones(2), np.ones(2), np.ones(2), np.ones(2), 1.0, 1.0).shape == (0,)


@pytest.mark.parametrize("T,S,k", [(1, 1, 1), (511, 3, 7), (513, 2, 50), (100_000, 4, 50), (2049, 64, 50)])
def test_rff_eval_vs_oracle_shapes(ctx, T, S, k):
    rng = np.random.default_rng(T + S + k)
    time = np.sort(rng.uniform(-20.0, 120.0, T))
    omega = rng.normal(size=(S, 2 * k)) * np.where(np.arange(2 * k) < k, 0.01, 1.0)
    a = rng.normal(size=(S, 2 * k)) * 0.15
    b = rng.normal(size=(S, 2 * k)) * 0.15
    shift = rng.normal(0, 1.0, S)
    amp = np.exp(rng.normal(size=S))
    got = ctx.rff_eval(time, omega, a, b, shift, amp)
    for s in range(S):
        exp = amp[s] * np.exp(orc.log_intensity_folded(time - shift[s], omega[s], a[s], b[s]))
        np.testing.assert_allclose(got[s], exp, rtol=RTOL)


def test_rff_eval_huge_phases(ctx):
    """Sky-scale shifts (1e3 s) and large absolute times: phases ~1e5 rad must still meet 1e-6."""
    rng = np.random.default_rng(9)
    time = 5.0e4 + np.linspace(0, 100, 5000)
    omega = rng.normal(size=(2, 100)) * 2.0
    a = rng.normal(size=(2, 100)) * 0.1
    b = rng.normal(size=(2, 100)) * 0.1
    shift = np.array([-1.0e3, 777.7])
    got = ctx.rff_eval(time, omega, a, b, shift, np.ones(2))
    for s in range(2):
        np.testing.assert_allclose(got[s], np.exp(orc.log_intensity_folded(time - shift[s], omega[s], a[s], b[s])), rtol=RTOL)


def test_rff_eval_phases_beyond_fp32_reduction(ctx):
    """|omega tau| / 2 pi >= 2^22 turns: the kernel must switch to its fp64 side path (per thread), not lose the phase."""
    rng = np.random.default_rng(10)
    time = np.concatenate([np.linspace(0.0, 10.0, 700), 6.0e4 + np.linspace(0, 1, 700)])  # small and huge phases mixed
    omega = rng.normal(size=(1, 20)) * 800.0
    a = rng.normal(size=(1, 20)) * 0.1
    b = rng.normal(size=(1, 20)) * 0.1
    got = ctx.rff_eval(time, omega, a, b, np.zeros(1), np.ones(1))
    assert np.abs(omega).max() * 6.0e4 / (2 * np.pi) > 2**22
    np.testing.assert_allclose(got[0], np.exp(orc.log_intensity_folded(time, omega[0], a[0], b[0])), rtol=RTOL)


@pytest.mark.parametrize("scale", [0.15, 0.45, 1.5])
def test_rff_eval_accuracy_tiers_by_amplitude(ctx, scale):
    """The error of the intensity is absolute in log(lambda) and scales with the draw's amplitudes; the kernel switches to
    compensated fp32 terms (sum a^2 + b^2 > 12) and to fp64 (> 100) so that 1e-6 relative holds for all of them."""
    rng = np.random.default_rng(int(scale * 100))
    T, S, J = 20_000, 3, 100
    time = np.sort(rng.uniform(0.0, 100.0, T))
    omega = rng.normal(size=(S, J)) * np.where(np.arange(J) < J // 2, 0.05, 2.0)
    a = rng.normal(size=(S, J)) * scale
    b = rng.normal(size=(S, J)) * scale
    a2 = (a * a + b * b).sum(axis=1)
    assert (a2.max() <= 12) if scale == 0.15 else ((12 < a2.min() and a2.max() <= 100) if scale == 0.45 else a2.min() > 100)
    shift = rng.normal(0, 0.5, S)
    got = ctx.rff_eval(time, omega, a, b, shift, np.ones(S))
    for s in range(S):
        logf = orc.log_intensity_folded(time - shift[s], omega[s], a[s], b[s])
        keep = np.abs(logf) < 600.0  # exp overflows float64 beyond (both sides give inf there)
        np.testing.assert_allclose(got[s][keep], np.exp(logf[keep]), rtol=RTOL)


def test_argument_errors(ctx):
    from pyipn_b200 import IpnError

    with pytest.raises(IpnError):
        ctx.rff_eval(np.arange(4.0), np.ones((1, 600)), np.ones((1, 600)), np.ones((1, 600)), [0.0], [1.0])  # J > IPN_MAX_J
    with pytest.raises(ValueError):
        ctx.rff_eval(np.arange(4.0), np.ones((2, 4)), np.ones((1, 4)), np.ones((2, 4)), [0.0, 0.0], [1.0, 1.0])


@pytest.mark.parametrize("T,S,k,scale", [(1024, 2, 50, 0.15), (100_000, 5, 50, 0.15), (4321, 3, 7, 0.45), (20_000, 4, 50, 1.5),
                                         (99_999, 2, 54, 0.2)])
def test_rff_eval_uniform_bins_take_the_contraction(ctx, T, S, k, scale):
    """Bin mid-points of np.arange edges (the only way the reference calls RFF: lightcurve.py:24-47 -> universe.py:441-529):
    ipn_rff_eval routes to the tcgen05 contraction (bin i = 128 p + q, one (T/128 x 2J).(2J x 128) product per draw, store
    epilogue).  Same tolerance, 1e-6 relative on lambda; per-draw shifts, negative and zero amplitudes, T not a multiple of
    128 or 32, large amplitudes (scale 1.5: the exponent bound 2^120 does not hold -> those draws fall to the direct kernel),
    and the forced direct kernel on the same inputs."""
    from pyipn_b200._lib import OPT_RFF_IMPL

    rng = np.random.default_rng(T + S + k)
    edges = -10.0 + 1e-3 * np.arange(T + 1)
    time = edges[:-1] + 0.5e-3
    J = 2 * k
    omega = rng.normal(size=(S, J)) * np.where(np.arange(J) < k, 0.01, 1.0)
    a = rng.normal(size=(S, J)) * scale
    b = rng.normal(size=(S, J)) * scale
    shift = rng.normal(0, 1.0, S)
    amp = np.exp(rng.normal(size=S))
    amp[0] = -amp[0]
    if S > 2:
        amp[2] = 0.0
    got = ctx.rff_eval(time, omega, a, b, shift, amp)
    assert ctx.last_rff_impl == 2
    for s in range(S):
        logf = orc.log_intensity_folded(time - shift[s], omega[s], a[s], b[s])
        keep = np.abs(logf) < 600.0
        np.testing.assert_allclose(got[s][keep], amp[s] * np.exp(logf[keep]), rtol=RTOL)
        if amp[s] == 0.0:
            assert np.all(got[s] == 0.0)
    ctx.set_option(OPT_RFF_IMPL, 1)
    try:
        direct = ctx.rff_eval(time, omega, a, b, shift, amp)
        assert ctx.last_rff_impl == 1
    finally:
        ctx.set_option(OPT_RFF_IMPL, 0)
    keep = np.isfinite(direct) & np.isfinite(got)
    np.testing.assert_allclose(got[keep], direct[keep], rtol=2 * RTOL)
    assert np.array_equal(np.isfinite(direct), np.isfinite(got))


def test_rff_eval_route_follows_the_bins(ctx):
    """Non-uniform bins, short light curves and J > 108 keep the direct kernel; the device-pointer entry point finds out on
    the device (one 24-byte read-back) and gives the same numbers as the host entry point."""
    import torch

    rng = np.random.default_rng(3)
    T, S, J = 5000, 3, 100
    time = 2.0 + 0.01 * (np.arange(T) + 0.5)
    omega = rng.normal(size=(S, J))
    a = rng.normal(size=(S, J)) * 0.1
    b = rng.normal(size=(S, J)) * 0.1
    shift, amp = rng.normal(size=S), np.exp(rng.normal(size=S))
    ref = ctx.rff_eval(time, omega, a, b, shift, amp)
    assert ctx.last_rff_impl == 2
    jig = time.copy()
    jig[1234] += 1e-7
    ctx.rff_eval(jig, omega, a, b, shift, amp)
    assert ctx.last_rff_impl == 1
    ctx.rff_eval(time[:500], omega, a, b, shift, amp)
    assert ctx.last_rff_impl == 1
    ctx.rff_eval(time, np.ones((1, 240)), np.zeros((1, 240)), np.zeros((1, 240)), [0.0], [1.0])
    assert ctx.last_rff_impl == 1
    dev = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in dict(time=time, omega=omega, a=a, b=b, shift=shift, amp=amp).items()}
    out = torch.empty((S, T), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    ctx.rff_eval_dev(T, dev["time"].data_ptr(), S, J, dev["omega"].data_ptr(), dev["a"].data_ptr(), dev["b"].data_ptr(),
                     dev["shift"].data_ptr(), dev["amp"].data_ptr(), out.data_ptr())
    ctx.synchronize()
    assert ctx.last_rff_impl == 2
    np.testing.assert_array_equal(out.cpu().numpy(), ref)


@pytest.mark.parametrize("J", [1, 3, 101, 107])
@pytest.mark.parametrize("uniform", [False, True])
def test_rff_eval_odd_number_of_frequencies(ctx, J, uniform):
    """J odd: the direct kernel's float4 parameter table must stay 16-byte aligned behind an odd number of doubles (it was
    not: misaligned LDS.128, a sticky fault), and the contraction pads 2J + 8 rows up to a multiple of 32 whatever J is."""
    rng = np.random.default_rng(J)
    T, S = 3000, 3
    time = 0.5e-3 + 1e-3 * np.arange(T) if uniform else np.sort(rng.uniform(0.0, 3.0, T))
    omega = rng.normal(size=(S, J)) * 3.0
    a = rng.normal(size=(S, J)) * 0.2
    b = rng.normal(size=(S, J)) * 0.2
    shift, amp = rng.normal(size=S), np.exp(rng.normal(size=S))
    got = ctx.rff_eval(time, omega, a, b, shift, amp)
    assert ctx.last_rff_impl == (2 if uniform else 1)
    for s in range(S):
        np.testing.assert_allclose(got[s], amp[s] * np.exp(orc.log_intensity_folded(time - shift[s], omega[s], a[s], b[s])), rtol=RTOL)


def test_rff_eval_contraction_across_draw_batches(ctx):
    """The contraction route works through the draws in internal batches like the likelihood grid: with IPN_OPT_DRAW_BATCH = 2 and
    five draws -- one of them with amplitudes beyond the 2^120 exponent bound, which the direct kernel takes over inside ITS
    batch -- every draw must match the oracle and the un-batched call."""
    from pyipn_b200._lib import OPT_DRAW_BATCH

    rng = np.random.default_rng(11)
    T, S, J = 4099, 5, 100
    time = 3.0 + 0.002 * (np.arange(T) + 0.5)
    omega = rng.normal(size=(S, J)) * 1.5
    a = rng.normal(size=(S, J)) * 0.12
    b = rng.normal(size=(S, J)) * 0.12
    a[3] *= 12.0  # sum_j |(a_j, b_j)| log2e > 120: unbounded exponent -> direct kernel for this draw only
    b[3] *= 12.0
    shift, amp = rng.normal(size=S), np.exp(rng.normal(size=S))
    whole = ctx.rff_eval(time, omega, a, b, shift, amp)
    assert ctx.last_rff_impl == 2
    ctx.set_option(OPT_DRAW_BATCH, 2)
    try:
        got = ctx.rff_eval(time, omega, a, b, shift, amp)
        assert ctx.last_rff_impl == 2 and ctx.last_draw_batch == 2
    finally:
        ctx.set_option(OPT_DRAW_BATCH, 0)
    np.testing.assert_array_equal(got, whole)
    for s in range(S):
        logf = orc.log_intensity_folded(time - shift[s], omega[s], a[s], b[s])
        keep = np.abs(logf) < 600.0
        np.testing.assert_allclose(got[s][keep], amp[s] * np.exp(logf[keep]), rtol=RTOL)
    assert (1.44269504 * np.sqrt(a[3] ** 2 + b[3] ** 2).sum()) > 120.0
