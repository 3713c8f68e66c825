"""ctypes access to the plain-C oracle (oracle/ipn_oracle.c).  TEST / BENCH INFRASTRUCTURE ONLY, like ipn_oracle.py."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "ipn_oracle.c")
LIB = os.path.join(HERE, "_build", "libipn_oracle.so")
_lib = None
_p, _i64, _f64 = C.c_void_p, C.c_int64, C.c_double


def build(force: bool = False) -> str:
    """gcc -O2 -pthread; rebuilt when the source is newer than the library."""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = [os.environ.get("CC", "gcc"), "-O2", "-fPIC", "-shared", "-pthread", "-o", LIB, SRC, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building the C oracle failed:\n" + r.stdout + r.stderr)
    return LIB


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.orc_max_threads.restype = C.c_int
        _lib.orc_set_threads.argtypes = [C.c_int]
        _lib.orc_set_threads.restype = None
        _lib.orc_loglik_folded.restype = _f64
        _lib.orc_loglik_folded.argtypes = [_i64, _p, _p, _p, _i64, _p, _p, _p, _f64, _f64, _f64]
        _lib.orc_loglik_delay_grid.argtypes = [_i64, _p, _p, _p, _i64, _i64, _p, _p, _p, _p, _p, _i64, _p, _p]
        _lib.orc_loglik_delay_grid.restype = None
        _lib.orc_loglik_sky_grid.argtypes = [_i64, _p, _i64, _p, _p, _p, _p, _i64, _p, _i64, _i64, _p, _p, _p, _p, _p, _p]
        _lib.orc_loglik_sky_grid.restype = None
        _lib.orc_expected_rate.argtypes = [_i64, _p, _i64, _i64, _p, _p, _p, _p, _p, _p, C.c_int, _p, _p, _i64, _p]
        _lib.orc_expected_rate.restype = None
    return _lib


def _d(x):
    return np.ascontiguousarray(x, dtype=np.float64)


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


def max_threads() -> int:
    return int(lib().orc_max_threads())


def set_threads(n: int) -> None:
    """0 = all online cores."""
    lib().orc_set_threads(int(n))


def loglik_delay_grid(counts, time, exposure, omega, a, b, amp, bkg, delays):
    """Same signature and result as ipn_oracle.loglik_delay_grid: ll (G, S)."""
    time, exposure = _d(time), _d(exposure)
    counts = np.ascontiguousarray(counts, dtype=np.int64)
    omega, a, b = np.atleast_2d(_d(omega)), np.atleast_2d(_d(a)), np.atleast_2d(_d(b))
    S, J = omega.shape
    amp, bkg, delays = _d(np.atleast_1d(amp)), _d(np.atleast_1d(bkg)), _d(np.atleast_1d(delays))
    ll = np.empty((delays.size, S))
    lib().orc_loglik_delay_grid(time.size, _ptr(time), _ptr(exposure), _ptr(counts), S, J, _ptr(omega), _ptr(a), _ptr(b),
                                _ptr(amp), _ptr(bkg), delays.size, _ptr(delays), _ptr(ll))
    return ll


def loglik_sky_grid(counts, time, exposure, n_time_bins, sc_pos, ghat, omega, a, b, amp, bkg):
    """Same signature and result as ipn_oracle.loglik_sky_grid: ll (P, S)."""
    time, exposure = np.atleast_2d(_d(time)), np.atleast_2d(_d(exposure))
    counts = np.atleast_2d(np.ascontiguousarray(counts, dtype=np.int64))
    D, Tmax = time.shape
    nb = np.ascontiguousarray(n_time_bins, dtype=np.int64)
    sc_pos, ghat = _d(sc_pos), np.atleast_2d(_d(ghat))
    omega, a, b = np.atleast_2d(_d(omega)), np.atleast_2d(_d(a)), np.atleast_2d(_d(b))
    S, J = omega.shape
    amp, bkg = np.atleast_2d(_d(amp)), np.atleast_2d(_d(bkg))
    ll = np.empty((ghat.shape[0], S))
    lib().orc_loglik_sky_grid(D, _ptr(nb), Tmax, _ptr(time), _ptr(exposure), _ptr(counts), _ptr(sc_pos), ghat.shape[0],
                              _ptr(ghat), S, J, _ptr(omega), _ptr(a), _ptr(b), _ptr(amp), _ptr(bkg), _ptr(ll))
    return ll


def expected_rate(time, omega1, omega2, beta1, beta2, bw, scale, amplitude, dt, N=None):
    """fit.py:651-691 (`_expeced_rate` for scale (S,), `_expeced_rate_multiscale` for scale (2, S)): out (N, T)."""
    time = _d(time)
    omega1, omega2, beta1, beta2 = (np.atleast_2d(_d(x)) for x in (omega1, omega2, beta1, beta2))
    k, S = omega1.shape
    scale = _d(scale)
    multi = 1 if scale.ndim == 2 else 0
    bw, amplitude, dt = _d(bw), _d(amplitude), _d(dt)
    N = S if N is None else int(N)
    out = np.empty((N, time.size))
    lib().orc_expected_rate(time.size, _ptr(time), k, S, _ptr(omega1), _ptr(omega2), _ptr(beta1), _ptr(beta2), _ptr(bw),
                            _ptr(scale), multi, _ptr(amplitude), _ptr(dt), N, _ptr(out))
    return out
