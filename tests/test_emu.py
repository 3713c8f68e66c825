"""CPU: the arithmetic of the tcgen05 delay-grid kernel, compiled from the kernel's own headers with g++
(tests/emu/tc_emu.cpp, IPN_HOST_EMU) and run at full size (1e5 bins) against the float64 oracle.

What this pins without a GPU: the digit-plane contraction is exact, the three epilogues (plain fp32, compensated fp32,
fp64) meet the 1e-3 nat tolerance of BASELINE.json at BASELINE config 2's shape, the result does not depend on the error
of the MUFU.LG2 seed, and the arg-max over the delays is the oracle's.  The GPU tests
(tests/test_gpu_loglik.py) check the same numbers on the device; agreement of the two is the evidence that what runs on
the tensor cores is this arithmetic and nothing else (tools/emu_vs_device.py measured it on a B200: 5.7e-8 nats apart on
every delay with the fp64 epilogue, profiles/r01_emu_vs_device.json).  The second half does the same for rff_eval_kernel's
per-term arithmetic (rff_terms.cuh) against the fixtures minted from the reference's own rff.py.
"""
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import c_oracle as corc
from pyipn_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "tc_emu.cpp")
BUILD = os.path.join(HERE, "emu", "_build")
EXE = os.path.join(BUILD, "tc_emu")
HEADERS = [os.path.join(HERE, "..", "pyipn_b200", "csrc", h) for h in ("loglik_common.cuh", "loglik_epilogue.cuh", "rff_terms.cuh")]

TOL_NATS = 1e-3  # BASELINE.json north_star: absolute 1e-3 nats on the log-likelihood sums


@pytest.fixture(scope="module")
def emu():
    os.makedirs(BUILD, exist_ok=True)
    newest = max(os.path.getmtime(p) for p in [SRC, *HEADERS])
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < newest:
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-o", EXE, SRC], check=True)
    corc.build()
    return EXE


def _run_grid(exe, tmp_path, w, s, delays, bias=4e-8, noise=6e-8, amp=None, counts=None):
    counts = w["counts"] if counts is None else counts
    amp = float(w["amp"][s]) if amp is None else amp
    path = os.path.join(tmp_path, "in.bin")
    with open(path, "wb") as f:
        np.array([w["time"].size, w["omega"].shape[1], delays.size], dtype=np.int64).tofile(f)
        np.array([amp, float(w["bkg"][s])], dtype=np.float64).tofile(f)
        for arr, dt in ((w["time"], np.float64), (w["exposure"], np.float64), (counts, np.int64), (w["omega"][s], np.float64),
                        (w["a"][s], np.float64), (w["b"][s], np.float64), (delays, np.float64)):
            np.ascontiguousarray(arr, dtype=dt).tofile(f)
    cmd = [exe, "grid", path, repr(bias), repr(noise)]
    out = subprocess.run(cmd, check=True, capture_output=True, text=True).stdout
    res = json.loads(out)
    ref = corc.loglik_delay_grid(counts, w["time"], w["exposure"], w["omega"][s:s + 1], w["a"][s:s + 1], w["b"][s:s + 1],
                                 [amp], w["bkg"][s:s + 1], delays)[:, 0]
    return res, np.array(res["ll"]), ref


def test_function_error_statistics(emu):
    """The numbers DESIGN.md section 3 quotes for the building blocks (2e6 samples each)."""
    st = json.loads(subprocess.run([emu, "funcs"], check=True, capture_output=True, text=True).stdout)
    # 2^x: no systematic error (a bias b costs b * sum kappa u ~ 1e5 b nats)
    for key in ("exp2_poly_rel", "exp2_hilo_rel"):
        assert abs(st[key]["mean"]) < 3e-10 and st[key]["rms"] < 3e-8 and st[key]["max"] < 1.2e-7, (key, st[key])
    # sin / cos of phases up to 1e5 rad, fp64-reduced and double-float-reduced
    for key in ("sin_phase_abs", "cos_phase_abs", "sin_phase_df_abs", "cos_phase_df_abs"):
        assert abs(st[key]["mean"]) < 2e-9 and st[key]["rms"] < 2.1e-8 and st[key]["max"] < 1.2e-7, (key, st[key])
    # log2(1 + u) = y0 + correction: unbiased whatever the seed's bias, and exact for u below half an ulp of 1
    for key in ("log2_1pu_abs_mufu_like", "log2_1pu_abs_bad_seed"):
        assert abs(st[key]["mean"]) < 5e-10 and st[key]["rms"] < 4e-8, (key, st[key])
    assert st["log2_1pu_tiny_rel"]["max"] < 2e-7, st["log2_1pu_tiny_rel"]
    # round 2: sine / cosine of the Phi image directly in 2^-30 fixed point (1024-point table + first-order rotation), phases
    # up to 1e5 rad: fifty times closer than the fp32 polynomial route
    for key in ("sin_phase_q30_abs", "cos_phase_q30_abs"):
        assert abs(st[key]["mean"]) < 2e-11 and st[key]["rms"] < 5e-10 and st[key]["max"] < 3e-9, (key, st[key])
    # integer exponent assembly + 2^x from (fraction, integer part): unbiased like exp2_poly, and the hi + lo form of the
    # compensated epilogue at less than half its random error
    assert abs(st["exp2_parts_rel"]["mean"]) < 3e-10 and st["exp2_parts_rel"]["rms"] < 3.2e-8 and st["exp2_parts_rel"]["max"] < 1.3e-7
    assert abs(st["exp2_parts_df_rel"]["mean"]) < 3e-10 and st["exp2_parts_df_rel"]["rms"] < 1.6e-8, st["exp2_parts_df_rel"]
    # the packed digit split of the image kernel is the one-at-a-time split; the interleaved bin order is a permutation
    assert st["digit_pack_mismatches"] == 0 and st["bin_position_errors"] == 0


def test_config2_full_size_all_epilogues(emu, tmp_path):
    """BASELINE config 2 shape (1e5 bins of 1 ms, K = 50): delays around the maximum and across the grid."""
    w = synth.delay_grid_workload(T=100_000, G=4096, S=1)
    g_true = int(np.argmin(np.abs(w["delays"] - w["true_delay"])))
    idx = np.unique(np.concatenate([np.arange(g_true - 4, g_true + 5), [0, 511, 1777, 3000, 4095]]))
    res, ll, ref = _run_grid(emu, str(tmp_path), w, 0, w["delays"][idx])
    assert res["Kpad"] == 224 and res["auto_mode"] == 1  # ~1 count per bin, sum n^2 = 4.5e5: the compensated fp32 epilogue
    err = ll - ref[:, None]
    assert np.abs(err[:, 0]).max() < TOL_NATS, err[:, 0]        # plain fp32
    assert np.abs(err[:, 1]).max() < TOL_NATS, err[:, 1]        # compensated fp32
    assert np.abs(err[:, 2]).max() < 2e-4, err[:, 2]            # fp64 epilogue: only the Phi quantisation is left
    for m in range(3):
        assert idx[np.argmax(ll[:, m])] == idx[np.argmax(ref)]
    # margin actually observed at this shape (rms ~1.5e-4 on the device): keep a regression guard well inside 1e-3
    assert np.sqrt(np.mean(err[:, 0] ** 2)) < 4e-4


def test_result_does_not_depend_on_the_lg2_seed_error(emu, tmp_path):
    """A seed with five times MUFU.LG2's measured bias, and a correctly rounded one, give the same sums (the refinement
    in poisson_term removes the seed's error; without it +4e-8 per bin would be ~3e-3 nats here)."""
    w = synth.delay_grid_workload(T=100_000, G=4096, S=1)
    g_true = int(np.argmin(np.abs(w["delays"] - w["true_delay"])))
    d = w["delays"][[g_true, 100]]
    _, ll_a, ref = _run_grid(emu, str(tmp_path), w, 0, d, bias=0.0, noise=0.0)
    _, ll_b, _ = _run_grid(emu, str(tmp_path), w, 0, d, bias=-2e-7, noise=3e-7)
    assert np.abs(ll_a[:, :2] - ll_b[:, :2]).max() < 3e-4
    assert np.abs(ll_b[:, :2] - ref[:, None]).max() < TOL_NATS


@pytest.mark.parametrize("seed", [1, 2, 5, 6])
def test_other_prior_draws_with_the_automatic_mode(emu, tmp_path, seed):
    """Other draws of the prior (seed 1: up to 113 counts per bin -> compensated fp32; seed 6: 1.1e6 counts, up to 406 per
    bin -> fp64 epilogue): the epilogue tc_draw_kernel picks holds 1e-3 nats wherever the model is not grossly wrong."""
    w = synth.delay_grid_workload(T=100_000, G=4096, S=1, seed=seed)
    g_true = int(np.argmin(np.abs(w["delays"] - w["true_delay"])))
    idx = np.array([g_true - 20, g_true - 1, g_true, g_true + 1, g_true + 20])
    res, ll, ref = _run_grid(emu, str(tmp_path), w, 0, w["delays"][idx])
    err = ll[:, res["auto_mode"]] - ref
    assert np.abs(err).max() < TOL_NATS, (res["auto_mode"], err)
    assert np.argmax(ll[:, res["auto_mode"]]) == np.argmax(ref)


def test_bright_data_mode_selection_and_accuracy(emu, tmp_path):
    """Tens to hundreds of counts per bin: the per-draw mode switch of tc_draw_kernel picks the epilogue that holds 1e-3."""
    for boost, bkg, expect_mode in ((400.0, 2.0e4, 1), (400.0, 4.0e5, 2)):
        w = synth.delay_grid_workload(T=20_000, G=64, S=1, boost=boost, bkg=bkg)
        g_true = int(np.argmin(np.abs(w["delays"] - w["true_delay"])))
        res, ll, ref = _run_grid(emu, str(tmp_path), w, 0, w["delays"][[g_true, 3]])
        assert res["auto_mode"] == expect_mode, res
        assert np.abs(ll[:, expect_mode] - ref).max() < TOL_NATS, (boost, bkg, ll - ref[:, None])


def test_edge_cases(emu, tmp_path):
    """amp = 0 (background only: exactly C_s), zero counts everywhere, ragged T (not a multiple of the 32-bin tile)."""
    w = synth.delay_grid_workload(T=1000 + 17, G=8, S=1)
    d = w["delays"][[0, 5]]
    _, ll, ref = _run_grid(emu, str(tmp_path), w, 0, d)
    assert np.abs(ll - ref[:, None]).max() < 1e-4
    _, ll0, ref0 = _run_grid(emu, str(tmp_path), w, 0, d, amp=0.0)
    assert np.abs(ll0 - ref0[:, None]).max() < 1e-9 * np.abs(ref0).max()
    zero = np.zeros_like(w["counts"])
    _, llz, refz = _run_grid(emu, str(tmp_path), w, 0, d, counts=zero)
    assert np.abs(llz - refz[:, None]).max() < 1e-4


def test_random_ragged_shapes(emu, tmp_path):
    """Fixed set of random small problems (T = 1..400 not a multiple of anything, J = 1..108, zero exposures, amp / bkg over
    eight decades, amp = 0, counts drawn from the model, from junk, or all zero): the epilogue the device would pick agrees
    with the float64 oracle to 1e-3 nats where the trial delay is the one the counts were drawn at, and to 1e-3 + 1e-6 |ll|
    for the junk (grossly wrong models of bright data, ll of -1e4 ... -1e7 nats: DESIGN.md section 3, last paragraph); -inf (counts in a bin of zero exposure) is reproduced exactly."""
    rng = np.random.default_rng(20261020)
    n_inf = 0
    for it in range(60):
        T, J, G = int(rng.integers(1, 400)), int(rng.integers(1, 109)), int(rng.integers(1, 4))
        span = 10 ** rng.uniform(-1, 3)
        time = np.sort(rng.uniform(0, span, T)) + rng.choice([0.0, 1e3, -5e2])
        expo = np.full(T, span / T) * rng.uniform(0.5, 1.5, T)
        if rng.random() < 0.3:
            expo[rng.integers(0, T, max(1, T // 10))] = 0.0
        bkg = 10 ** rng.uniform(-1, 5)
        amp = 0.0 if rng.random() < 0.1 else bkg * 10 ** rng.uniform(-6, 2)
        a, b = rng.normal(size=J), rng.normal(size=J)
        sc = np.sqrt(10 ** rng.uniform(-2, 1.3) / (a ** 2 + b ** 2).sum())
        a, b = a * sc, b * sc
        omega = rng.normal(size=J) * 10 ** rng.uniform(-2, 2)
        delays = np.concatenate([[0.0], rng.uniform(-1, 1, G - 1) * 10 ** rng.uniform(-3, 1)])
        mu = expo * (amp * np.exp(np.cos(np.outer(time, omega)) @ a + np.sin(np.outer(time, omega)) @ b) + bkg)
        kind = it % 3
        counts = rng.poisson(np.minimum(mu, 1e6)) if kind == 0 else (rng.integers(0, 5, T) if kind == 1 else np.zeros(T))
        counts = np.asarray(counts, dtype=np.int64)
        if rng.random() < 0.7:
            counts[expo == 0] = 0
        w = dict(time=time, exposure=expo, counts=counts, omega=omega[None], a=a[None], b=b[None], amp=np.array([amp]),
                 bkg=np.array([bkg]), delays=delays)
        res, ll, ref = _run_grid(emu, str(tmp_path), w, 0, delays)
        got = ll[:, res["auto_mode"]]
        inf = np.isneginf(ref)
        n_inf += int(inf.sum())
        assert np.array_equal(np.isneginf(got), inf) and not np.isnan(got).any(), (it, got, ref)
        err = np.abs(got[~inf] - ref[~inf])
        assert (err < 1e-3 + 1e-6 * np.abs(ref[~inf])).all(), (it, T, J, res["auto_mode"], got, ref)
        if kind == 0 and not inf[0] and mu.max() < 1e6:
            assert err[0] < TOL_NATS, (it, T, J, res["auto_mode"], got[0], ref[0])
    assert n_inf > 0  # the -inf branch was exercised


# ---------------------------------------------------------------------------------------------------------------
# rff_eval_kernel's arithmetic (pyipn_b200/csrc/rff_terms.cuh) against the fixtures minted from the reference's rff.py
# ---------------------------------------------------------------------------------------------------------------
RTOL_LAMBDA = 1e-6  # BASELINE.json north_star: relative 1e-6 on lambda


def _run_rff(exe, tmp_path, time, omega, a, b, shift, amp):
    time = np.ascontiguousarray(time, dtype=np.float64)
    omega, a, b = (np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64) for x in (omega, a, b))
    S, J = omega.shape
    src, dst = os.path.join(tmp_path, "rff_in.bin"), os.path.join(tmp_path, "rff_out.bin")
    with open(src, "wb") as f:
        np.array([time.size, S, J], dtype=np.int64).tofile(f)
        for arr in (time, omega, a, b, np.broadcast_to(np.asarray(shift, dtype=np.float64), (S,)),
                    np.broadcast_to(np.asarray(amp, dtype=np.float64), (S,))):
            np.ascontiguousarray(arr, dtype=np.float64).tofile(f)
    subprocess.run([exe, "rff", src, dst], check=True)
    raw = np.fromfile(dst, dtype=np.float64)
    return raw[:S * time.size].reshape(S, time.size), raw[S * time.size:].astype(int)


def test_rff_arithmetic_on_the_reference_fixtures(emu, tmp_path, rff_golden):
    """RFF / RFF_multiscale (rff.py:5-29) and the strided draw loops (fit.py:651-691): the kernel's fp32 double-float phase
    reduction, minimax sin / cos and Kahan sum reproduce the reference's own outputs to 1e-6."""
    from pyipn_b200.params import fold_draws

    g = rff_golden
    for scale, want in ((g["A_scale"][1], g["A_rff"]), (g["A_scale"], g["A_rff_ms"])):
        sc = np.asarray(scale, dtype=float).reshape(-1, 1) if np.ndim(scale) else scale
        w, a, b = fold_draws(g["A_omega1"], g["A_omega2"], g["A_beta1"], g["A_beta2"], float(g["A_bw"]), sc)
        got, _ = _run_rff(emu, str(tmp_path), g["A_time"], w, a, b, 0.0, 1.0)
        np.testing.assert_allclose(got[0], want, rtol=RTOL_LAMBDA)
    for scale, want in ((g["B_scale"][1], g["B_rate_single"]), (g["B_scale"], g["B_rate_multi"])):
        w, a, b = fold_draws(g["B_omega1"], g["B_omega2"], g["B_beta1"], g["B_beta2"], g["B_bw"], scale)
        got, tiers = _run_rff(emu, str(tmp_path), g["B_time"], w, a, b, g["B_dt"], g["B_amp"])
        np.testing.assert_allclose(got, want, rtol=RTOL_LAMBDA)
    # one bin, k = 1, |phase| ~ 9e3 rad
    w, a, b = fold_draws(g["C_omega1"], g["C_omega2"], g["C_beta1"], g["C_beta2"], 1.0, np.array([[0.4], [1.3]]))
    got, _ = _run_rff(emu, str(tmp_path), g["C_time"], w, a, b, 0.0, 1.0)
    np.testing.assert_allclose(got[0], g["C_rff_ms"], rtol=RTOL_LAMBDA)


@pytest.mark.parametrize("scale,tier", [(0.15, 0), (0.45, 1), (1.5, 2)])
def test_rff_accuracy_tiers_and_huge_phases(emu, tmp_path, scale, tier):
    """Same cases as tests/test_gpu_rff.py: every amplitude tier, at |phase| up to 1e5 rad (sky-scale shift), holds 1e-6."""
    from oracle import ipn_oracle as orc

    rng = np.random.default_rng(int(scale * 100))
    T, S, J = 20_000, 2, 100
    time = 5.0e4 + np.sort(rng.uniform(0.0, 100.0, T))
    omega = rng.normal(size=(S, J)) * np.where(np.arange(J) < J // 2, 0.05, 2.0)
    a, b = rng.normal(size=(S, J)) * scale, rng.normal(size=(S, J)) * scale
    shift = np.array([-1.0e3, 777.7])
    got, tiers = _run_rff(emu, str(tmp_path), time, omega, a, b, shift, 1.0)
    assert (tiers == tier).all(), tiers
    for s in range(S):
        logf = orc.log_intensity_folded(time - shift[s], omega[s], a[s], b[s])
        keep = np.abs(logf) < 600.0
        np.testing.assert_allclose(got[s][keep], np.exp(logf[keep]), rtol=RTOL_LAMBDA)


def test_rff_phase_beyond_the_fp32_reduction_takes_the_fp64_path(emu, tmp_path):
    from oracle import ipn_oracle as orc

    rng = np.random.default_rng(10)
    time = np.concatenate([np.linspace(0.0, 10.0, 300), 6.0e4 + np.linspace(0, 1, 300)])
    omega = rng.normal(size=(1, 20)) * 800.0
    a, b = rng.normal(size=(1, 20)) * 0.1, rng.normal(size=(1, 20)) * 0.1
    got, tiers = _run_rff(emu, str(tmp_path), time, omega, a, b, 0.0, 1.0)
    assert tiers[0] == 2 and np.abs(omega).max() * 6.0e4 / (2 * np.pi) > 2 ** 22
    np.testing.assert_allclose(got[0], np.exp(orc.log_intensity_folded(time, omega[0], a[0], b[0])), rtol=RTOL_LAMBDA)


# ---------------------------------------------------------------------------------------------------------------
# the bin-permutation kernels themselves (pyipn_b200/csrc/tc_perm.cuh), compiled for the host with a CUDA-thread shim
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def perm_emu():
    src = os.path.join(HERE, "emu", "perm_emu.cpp")
    exe = os.path.join(BUILD, "perm_emu")
    hdrs = [os.path.join(HERE, "..", "pyipn_b200", "csrc", h) for h in ("tc_perm.cuh", "loglik_common.cuh")]
    os.makedirs(BUILD, exist_ok=True)
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(p) for p in [src, *hdrs]):
        subprocess.run(["g++", "-O1", "-std=c++20", "-pthread", "-o", exe, src], check=True)
    return exe


@pytest.mark.parametrize("T,seed,mean", [(0, 1, 0.9), (1, 2, 0.9), (31, 3, 0.9), (1024, 4, 0.9), (1025, 5, 2.0), (4100, 6, 0.8),
                                          (3333, 7, 0.1), (2500, 8, 6.0)])
def test_permutation_kernels_run_from_their_own_source(perm_emu, T, seed, mean):
    """tc_perm_kernel (bins with counts first): every CUDA thread an OS thread, real barriers and warp collectives; result =
    std::stable_sort, flags (bins with counts, mode, sum n^2) exact, nothing written past the end.  Empty input, partial
    blocks, several blocks, sparse and bright data."""
    r = subprocess.run([perm_emu, str(T), str(seed), str(mean)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr


@pytest.mark.parametrize("layout", [0, 2, 3])
def test_alternative_bin_layouts_are_permutations_too(layout, tmp_path):
    """The bin orders kept as build options (-DIPN_TC_INTERLEAVE=0 natural, 2 / 3 bins with and without counts mixed within
    every tile; measured slower, profiles/r02_tile_layouts_ab.txt) stay correct: the permutation kernel compiled with the
    option reproduces the order built tile by tile, for sparse, balanced and bright data and a partial last tile."""
    src = os.path.join(HERE, "emu", "perm_emu.cpp")
    exe = str(tmp_path / f"perm_emu_{layout}")
    subprocess.run(["g++", "-O1", "-std=c++20", "-pthread", f"-DIPN_TC_INTERLEAVE={layout}", "-o", exe, src], check=True)
    for T, seed, mean in [(0, 1, 0.9), (31, 3, 0.9), (1025, 5, 2.0), (4100, 6, 0.8), (3333, 7, 0.1), (2500, 8, 6.0), (64, 10, 0.7)]:
        r = subprocess.run([exe, str(T), str(seed), str(mean)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and r.stdout.startswith("ok"), (layout, T, r.stdout + r.stderr)


# ---------------------------------------------------------------------------------------------------------------
# the operand-image kernels themselves (tc_images.cuh + tc_perm.cuh) on the host, against the restatement used above
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def images_emu():
    src = os.path.join(HERE, "emu", "images_emu.cpp")
    exe = os.path.join(BUILD, "images_emu")
    deps = [src, os.path.join(HERE, "emu", "cuda_shim.h")] + [os.path.join(HERE, "..", "pyipn_b200", "csrc", h) for h in
                                                              ("tc_images.cuh", "tc_perm.cuh", "loglik_epilogue.cuh", "loglik_common.cuh")]
    os.makedirs(BUILD, exist_ok=True)
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(p) for p in deps):
        subprocess.run(["g++", "-O1", "-std=c++20", "-pthread", "-o", exe, src], check=True)
    return exe


@pytest.mark.parametrize("T,G,kw", [(1000 + 17, 130, {}), (333, 5, dict(bkg=15000.0, boost=12000.0)), (64, 1, {}),
                                     (700, 129, dict(bkg=1.0e6, boost=1.0e6))])
def test_restated_images_equal_the_kernels_own(emu, images_emu, tmp_path, T, G, kw):
    """tc_perm*_kernel, tc_draw_kernel, tc_wimg_kernel and tc_pimg_kernel run on the host from their own source (every CUDA
    thread an OS thread): scale, epilogue choice, W image, Phi image, per-bin constants and tile flags are byte for byte what
    the restatement inside tc_emu builds.  So the full-size emulation above runs on the images the device builds."""
    w = synth.delay_grid_workload(T=T, G=G, S=1, **kw)
    path = os.path.join(str(tmp_path), "in.bin")
    with open(path, "wb") as f:
        np.array([T, w["omega"].shape[1], G], dtype=np.int64).tofile(f)
        np.array([float(w["amp"][0]), float(w["bkg"][0])], dtype=np.float64).tofile(f)
        for arr, dt in ((w["time"], np.float64), (w["exposure"], np.float64), (w["counts"], np.int64), (w["omega"][0], np.float64),
                        (w["a"][0], np.float64), (w["b"][0], np.float64), (w["delays"], np.float64)):
            np.ascontiguousarray(arr, dtype=dt).tofile(f)
    a, b = os.path.join(str(tmp_path), "restated.bin"), os.path.join(str(tmp_path), "kernels.bin")
    subprocess.run([emu, "images", path, a], check=True)
    subprocess.run([images_emu, path, b], check=True, timeout=600)
    A, B = np.fromfile(a, dtype=np.uint8), np.fromfile(b, dtype=np.uint8)
    assert A.size == B.size and np.array_equal(A, B), (A.size, B.size, np.nonzero(A != B)[0][:8] if A.size == B.size else None)
    assert int(B[:16].view(np.int32)[3]) in (0, 1, 2)
