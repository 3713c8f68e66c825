"""ctypes binding of libipn_b200.so (the C ABI declared in include/ipn.h).

There is no CPU fallback: if the shared library is missing, or no B200 is visible, every compute entry
point raises ``IpnError`` -- loudly -- instead of computing anything on the host.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IPN_B200_LIB") or os.path.join(HERE, "libipn_b200.so")  # IPN_B200_LIB: explicit path to the library
# the same library built with -DIPN_TC_HOOKS (pacing / work-skipping switches of the hot kernel); tests/ and tools/ only
HOOKS_LIB_PATH = os.path.join(HERE, "libipn_b200_hooks.so")

IPN_OK = 0
IPN_ERR_INVALID_ARGUMENT = -1
IPN_ERR_NO_DEVICE = -2
IPN_ERR_CUDA = -3
IPN_ERR_ALLOC = -4
IPN_ERR_UNSUPPORTED = -5

OPT_LOGLIK_IMPL = 1
OPT_PRECISE_W = 2
OPT_DRAW_BATCH = 3
OPT_HIFI = 4
OPT_RFF_IMPL = 5

# every symbol include/ipn.h declares (tests/test_abi.py checks the list against the header and the .so)
EXPORTS = (
    "ipn_ctx_create", "ipn_ctx_destroy", "ipn_ctx_set_stream", "ipn_ctx_use_own_stream", "ipn_ctx_synchronize", "ipn_ctx_set_option",
    "ipn_ctx_launch_count", "ipn_ctx_last_kernel_ms", "ipn_ctx_last_main_kernel_ms", "ipn_ctx_last_loglik_impl", "ipn_ctx_last_rff_impl", "ipn_ctx_last_draw_batch", "ipn_last_error", "ipn_version",
    "ipn_rff_eval", "ipn_rff_eval_dev", "ipn_loglik_delay_grid", "ipn_loglik_delay_grid_dev",
    "ipn_loglik_sky_grid", "ipn_loglik_sky_grid_dev", "ipn_correlate", "ipn_bin_events", "ipn_bin_events_dev", "ipn_thinning_events", "ipn_poisson_counts", "ipn_poisson_counts_dev", "ipn_microbench", "ipn_sky_credible", "ipn_sky_credible_dev",
)


class IpnError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libipn_b200 error {code}: {message}")
        self.code = code


_libs = {}
_lib_lock = threading.Lock()

_p = C.c_void_p
_i64 = C.c_int64


def load_library(path: str | None = None):
    """Load (once per path) and return the ctypes handle; raises IpnError if the CUDA library has not been built."""
    path = path or LIB_PATH
    with _lib_lock:
        if path in _libs:
            return _libs[path]
        if not os.path.exists(path):
            raise IpnError(
                IPN_ERR_NO_DEVICE,
                f"{path} not found: build it with `python -m pyipn_b200._build` "
                "(pyipn_b200 has no CPU fallback and will not compute without its CUDA library)",
            )
        lib = C.CDLL(path)
        lib.ipn_last_error.restype = C.c_char_p
        lib.ipn_version.restype = C.c_char_p
        lib.ipn_ctx_create.argtypes = [C.c_int, C.POINTER(_p)]
        lib.ipn_ctx_destroy.argtypes = [_p]
        lib.ipn_ctx_destroy.restype = None
        lib.ipn_ctx_set_stream.argtypes = [_p, _p]
        lib.ipn_ctx_use_own_stream.argtypes = [_p]
        lib.ipn_ctx_synchronize.argtypes = [_p]
        lib.ipn_ctx_set_option.argtypes = [_p, C.c_int, _i64]
        lib.ipn_ctx_launch_count.argtypes = [_p]
        lib.ipn_ctx_launch_count.restype = C.c_uint64
        lib.ipn_ctx_last_kernel_ms.argtypes = [_p]
        lib.ipn_ctx_last_kernel_ms.restype = C.c_double
        lib.ipn_ctx_last_loglik_impl.argtypes = [_p]
        lib.ipn_ctx_last_rff_impl.argtypes = [_p]
        lib.ipn_ctx_last_draw_batch.argtypes = [_p]
        lib.ipn_ctx_last_draw_batch.restype = _i64
        lib.ipn_ctx_last_main_kernel_ms.argtypes = [_p, C.POINTER(_i64)]
        lib.ipn_ctx_last_main_kernel_ms.restype = C.c_double
        rff_args = [_p, _i64, _p, _i64, _i64, _p, _p, _p, _p, _p, _p]
        lib.ipn_rff_eval.argtypes = rff_args
        lib.ipn_rff_eval_dev.argtypes = rff_args
        grid_args = [_p, _i64, _p, _p, _p, _i64, _i64, _p, _p, _p, _p, _p, _i64, _p, _p]
        lib.ipn_loglik_delay_grid.argtypes = grid_args
        lib.ipn_loglik_delay_grid_dev.argtypes = grid_args
        lib.ipn_loglik_sky_grid.argtypes = [_p, _i64, _p, _i64, _p, _p, _p, _p, _i64, _p, _i64, _i64, _p, _p, _p, _p, _p, _p]
        lib.ipn_loglik_sky_grid_dev.argtypes = lib.ipn_loglik_sky_grid.argtypes
        pc_args = [_p, _i64, _p, _i64, _p, _p, C.c_uint64, _p]
        lib.ipn_poisson_counts.argtypes = pc_args
        lib.ipn_poisson_counts_dev.argtypes = pc_args
        lib.ipn_thinning_events.argtypes = [_p, C.c_double, C.c_double, _i64, _p, _p, _p, _p, C.c_double, C.c_double, C.c_uint64,
                                            _i64, _p, C.POINTER(_i64)]
        lib.ipn_bin_events.argtypes = [_p, _i64, _p, C.c_double, C.c_double, _i64, _p]
        lib.ipn_bin_events_dev.argtypes = [_p, _i64, _p, C.c_double, C.c_double, _i64, _p]
        lib.ipn_correlate.argtypes = [_p, _i64, _p, _p, _p, _i64, _p, _p, _p, _i64, _i64, _i64, _i64, C.c_double, C.c_double,
                                      C.c_double, _i64, _p, _p]
        lib.ipn_microbench.argtypes = [_p, C.c_int, C.POINTER(C.c_double)]
        sky_args = [_p, _i64, _i64, _p, _p, _p, _p, _p, _p]
        lib.ipn_sky_credible.argtypes = sky_args
        lib.ipn_sky_credible_dev.argtypes = sky_args
        _libs[path] = lib
        return lib


def _check(lib, rc):
    if rc != IPN_OK:
        raise IpnError(rc, lib.ipn_last_error().decode("utf-8", "replace"))


def _f64(x, shape=None, name="array"):
    arr = np.ascontiguousarray(x, dtype=np.float64)
    if shape is not None and tuple(arr.shape) != tuple(shape):
        raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(arr.shape)}")
    return arr


def _ptr(arr):
    return C.c_void_p(arr.ctypes.data)


class Context:
    """One libipn_b200 context = one GPU.  Host-array methods take/return numpy; ``*_dev`` take raw device pointers."""

    def __init__(self, device: int = 0, lib_path: str | None = None):
        self._lib = load_library(lib_path)
        h = _p()
        _check(self._lib, self._lib.ipn_ctx_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ipn_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- plumbing ------------------------------------------------------------------------------------
    def set_stream(self, cuda_stream: int | None):
        """Run on ``cuda_stream`` (a raw ``cudaStream_t`` handle such as ``torch.cuda.current_stream().cuda_stream``; 0 is
        CUDA's legacy default stream, taken literally); ``None`` returns to the context's own stream."""
        if cuda_stream is None:
            _check(self._lib, self._lib.ipn_ctx_use_own_stream(self._h))
        else:
            _check(self._lib, self._lib.ipn_ctx_set_stream(self._h, _p(int(cuda_stream))))

    def synchronize(self):
        _check(self._lib, self._lib.ipn_ctx_synchronize(self._h))

    def set_option(self, key: int, value: int):
        _check(self._lib, self._lib.ipn_ctx_set_option(self._h, int(key), int(value)))

    @property
    def launch_count(self) -> int:
        return int(self._lib.ipn_ctx_launch_count(self._h))

    @property
    def last_kernel_ms(self) -> float:
        return float(self._lib.ipn_ctx_last_kernel_ms(self._h))

    @property
    def last_main_kernel(self):
        """(summed ms, launches) of the dominant contraction kernel in the most recent compute call."""
        n = _i64(0)
        ms = float(self._lib.ipn_ctx_last_main_kernel_ms(self._h, C.byref(n)))
        return ms, int(n.value)

    @property
    def last_loglik_impl(self) -> int:
        return int(self._lib.ipn_ctx_last_loglik_impl(self._h))

    @property
    def last_rff_impl(self) -> int:
        """1 = direct kernel, 2 = tcgen05 contraction on uniform bins (the most recent ``rff_eval`` call)."""
        return int(self._lib.ipn_ctx_last_rff_impl(self._h))

    @property
    def last_draw_batch(self) -> int:
        return int(self._lib.ipn_ctx_last_draw_batch(self._h))

    def microbench(self, which: int) -> float:
        out = C.c_double(0.0)
        _check(self._lib, self._lib.ipn_microbench(self._h, int(which), C.byref(out)))
        return out.value

    # ---- host-array entry points -----------------------------------------------------------------------
    def rff_eval(self, time, omega, a, b, shift, amp):
        time = _f64(time).ravel()
        omega = np.atleast_2d(_f64(omega))
        S, J = omega.shape
        a = _f64(a, (S, J), "a") if np.ndim(a) == 2 else _f64(np.atleast_2d(a), (S, J), "a")
        b = _f64(b, (S, J), "b") if np.ndim(b) == 2 else _f64(np.atleast_2d(b), (S, J), "b")
        shift = _f64(np.atleast_1d(shift), (S,), "shift")
        amp = _f64(np.atleast_1d(amp), (S,), "amp")
        out = np.empty((S, time.size), dtype=np.float64)
        _check(self._lib, self._lib.ipn_rff_eval(self._h, time.size, _ptr(time), S, J, _ptr(omega), _ptr(a), _ptr(b),
                                                  _ptr(shift), _ptr(amp), _ptr(out)))
        return out

    def loglik_delay_grid(self, time, exposure, counts, omega, a, b, amp, bkg, delays, out=None):
        time = _f64(time).ravel()
        T = time.size
        exposure = _f64(exposure, (T,), "exposure")
        counts = np.ascontiguousarray(counts, dtype=np.int64)
        if counts.shape != (T,):
            raise ValueError(f"counts: expected shape {(T,)}, got {counts.shape}")
        omega = np.atleast_2d(_f64(omega))
        S, J = omega.shape
        a = _f64(np.atleast_2d(a), (S, J), "a")
        b = _f64(np.atleast_2d(b), (S, J), "b")
        amp = _f64(np.atleast_1d(amp), (S,), "amp")
        bkg = _f64(np.atleast_1d(bkg), (S,), "bkg")
        delays = _f64(np.atleast_1d(delays)).ravel()
        G = delays.size
        if out is None:
            out = np.empty((G, S), dtype=np.float64)
        elif out.shape != (G, S) or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float64 array of shape (G, S)")
        _check(self._lib, self._lib.ipn_loglik_delay_grid(
            self._h, T, _ptr(time), _ptr(exposure), _ptr(counts), S, J, _ptr(omega), _ptr(a), _ptr(b), _ptr(amp),
            _ptr(bkg), G, _ptr(delays), _ptr(out)))
        return out

    def loglik_sky_grid(self, nbins, time, exposure, counts, sc_pos, ghat, omega, a, b, amp, bkg):
        time = np.atleast_2d(_f64(time))
        D, Tmax = time.shape
        exposure = _f64(np.atleast_2d(exposure), (D, Tmax), "exposure")
        counts = np.ascontiguousarray(np.atleast_2d(counts), dtype=np.int64)
        if counts.shape != (D, Tmax):
            raise ValueError(f"counts: expected shape {(D, Tmax)}, got {counts.shape}")
        nbins = np.ascontiguousarray(nbins, dtype=np.int64)
        if nbins.shape != (D,):
            raise ValueError(f"nbins: expected shape {(D,)}, got {nbins.shape}")
        sc_pos = _f64(sc_pos, (D, 3), "sc_pos")
        ghat = np.atleast_2d(_f64(ghat))
        if ghat.shape[1] != 3:
            raise ValueError("ghat must be (P, 3)")
        P = ghat.shape[0]
        omega = np.atleast_2d(_f64(omega))
        S, J = omega.shape
        a = _f64(np.atleast_2d(a), (S, J), "a")
        b = _f64(np.atleast_2d(b), (S, J), "b")
        amp = _f64(np.atleast_2d(amp), (D, S), "amp")
        bkg = _f64(np.atleast_2d(bkg), (D, S), "bkg")
        out = np.empty((P, S), dtype=np.float64)
        _check(self._lib, self._lib.ipn_loglik_sky_grid(
            self._h, D, _ptr(nbins), Tmax, _ptr(time), _ptr(exposure), _ptr(counts), _ptr(sc_pos), P, _ptr(ghat), S, J,
            _ptr(omega), _ptr(a), _ptr(b), _ptr(amp), _ptr(bkg), _ptr(out)))
        return out

    def loglik_sky_grid_dev(self, nbins, Tmax, d_time, d_exposure, d_counts, d_sc_pos, P, d_ghat, S, J, d_omega, d_a, d_b, d_amp,
                            d_bkg, d_ll):
        """``loglik_sky_grid`` on device-resident arrays given as raw device pointers (ints); ``nbins`` is a host sequence of
        the D per-detector bin counts.  Asynchronous on the context's stream; ``d_ll`` (P x S float64) is overwritten."""
        nbins = np.ascontiguousarray(nbins, dtype=np.int64)
        _check(self._lib, self._lib.ipn_loglik_sky_grid_dev(
            self._h, int(nbins.size), _ptr(nbins), int(Tmax), _p(d_time), _p(d_exposure), _p(d_counts), _p(d_sc_pos), int(P),
            _p(d_ghat), int(S), int(J), _p(d_omega), _p(d_a), _p(d_b), _p(d_amp), _p(d_bkg), _p(d_ll)))

    def thinning_events(self, tstart, tstop, K=(), t_start=(), t_rise=(), t_decay=(), bkg_slope=0.0, bkg_intercept=0.0,
                        seed=1234):
        """Arrival times on [tstart, tstop] of sum-of-Norris-pulses + linear background (possion_gen.py:159-242)."""
        K, t_start, t_rise, t_decay = (_f64(np.atleast_1d(x)).ravel() for x in (K, t_start, t_rise, t_decay))
        n = K.size
        if not (t_start.size == t_rise.size == t_decay.size == n):
            raise ValueError("pulse parameter arrays must have the same length")
        # capacity from the integral bound fmax*(span) is only known on the device: start generous, grow on demand
        cap = 1 << 16
        cnt = _i64(0)
        while True:
            out = np.empty(cap, dtype=np.float64)
            rc = self._lib.ipn_thinning_events(self._h, float(tstart), float(tstop), n, _ptr(K), _ptr(t_start), _ptr(t_rise),
                                               _ptr(t_decay), float(bkg_slope), float(bkg_intercept),
                                               C.c_uint64(int(seed) & (2**64 - 1)), cap, _ptr(out), C.byref(cnt))
            if rc == IPN_ERR_INVALID_ARGUMENT and cnt.value > cap:
                cap = int(cnt.value) + 16
                continue
            _check(self._lib, rc)
            return out[: cnt.value].copy()

    def bin_events(self, times, tstart, tstop, dt):
        """(rate, edges, counts) with the semantics of LightCurve.get_binned_light_curve (lightcurve.py:24-47)."""
        edges = np.arange(tstart, tstop, dt)
        times = _f64(times).ravel()
        counts = np.zeros(max(len(edges) - 1, 0), dtype=np.int64)
        if len(edges) >= 2:
            # numpy fills a float arange as start + i*delta with delta = edges[1] - edges[0] (not dt itself)
            _check(self._lib, self._lib.ipn_bin_events(self._h, times.size, _ptr(times), float(edges[0]),
                                                        float(edges[1] - edges[0]), len(edges), _ptr(counts)))
        return counts / float(dt), edges, counts

    def correlate(self, t1, c1, e1, t2, c2, e2, i_beg_2, i_end_2, i_beg_1, n_max_1, fscale, dt_1, dt_2, n_sum_2=1):
        t1, c1, e1 = _f64(t1).ravel(), _f64(c1).ravel(), _f64(e1).ravel()
        t2, c2, e2 = _f64(t2).ravel(), _f64(c2).ravel(), _f64(e2).ravel()
        n1, n2 = c1.size, c2.size
        if e1.size != n1 or t1.size < n1 or e2.size != n2 or t2.size < n2:
            raise ValueError("light-curve arrays have inconsistent lengths")
        arr_dt = np.zeros(int(n_max_1))
        arr_chi = np.zeros(int(n_max_1))
        _check(self._lib, self._lib.ipn_correlate(self._h, n1, _ptr(t1), _ptr(c1), _ptr(e1), n2, _ptr(t2), _ptr(c2), _ptr(e2),
                                                   int(i_beg_2), int(i_end_2), int(i_beg_1), int(n_max_1), float(fscale),
                                                   float(dt_1), float(dt_2), int(n_sum_2), _ptr(arr_dt), _ptr(arr_chi)))
        return arr_dt, arr_chi

    def sky_credible(self, ll, log_weight=None):
        """Credible-region products of a per-pixel log-likelihood: ``ll`` is (P,) or (P, S) (draws are marginalised).

        Returns ``(prob, level, order, log_norm)``: normalised pixel probabilities, the cumulative probability at which
        each pixel enters the greedy (highest-density-first) region, the pixels in that order, and the log normaliser."""
        ll = _f64(ll)
        if ll.ndim == 1:
            ll = ll[:, None]
        if ll.ndim != 2:
            raise ValueError("ll must be (P,) or (P, S)")
        ll = np.ascontiguousarray(ll)
        P, S = ll.shape
        lw = None if log_weight is None else _f64(log_weight, (P,), "log_weight")
        prob, level = np.empty(P), np.empty(P)
        order = np.empty(P, dtype=np.int64)
        log_norm = np.empty(1)
        _check(self._lib, self._lib.ipn_sky_credible(self._h, P, S, _ptr(ll), None if lw is None else _ptr(lw), _ptr(prob),
                                                      _ptr(level), _ptr(order), _ptr(log_norm)))
        return prob, level, order, float(log_norm[0])

    def poisson_counts(self, exposure, rates, bkg, seed):
        rates = np.atleast_2d(_f64(rates))
        S, T = rates.shape
        exposure = _f64(exposure, (T,), "exposure")
        bkg = _f64(np.atleast_1d(bkg), (S,), "bkg")
        out = np.empty((S, T), dtype=np.int64)
        _check(self._lib, self._lib.ipn_poisson_counts(self._h, T, _ptr(exposure), S, _ptr(rates), _ptr(bkg),
                                                        C.c_uint64(int(seed) & (2**64 - 1)), _ptr(out)))
        return out

    # ---- device-pointer entry points (ints = CUdeviceptr, e.g. torch.Tensor.data_ptr()) -------------------
    def rff_eval_dev(self, T, d_time, S, J, d_omega, d_a, d_b, d_shift, d_amp, d_out):
        _check(self._lib, self._lib.ipn_rff_eval_dev(self._h, T, _p(d_time), S, J, _p(d_omega), _p(d_a), _p(d_b),
                                                      _p(d_shift), _p(d_amp), _p(d_out)))

    def loglik_delay_grid_dev(self, T, d_time, d_exposure, d_counts, S, J, d_omega, d_a, d_b, d_amp, d_bkg, G,
                              d_delays, d_ll):
        _check(self._lib, self._lib.ipn_loglik_delay_grid_dev(
            self._h, T, _p(d_time), _p(d_exposure), _p(d_counts), S, J, _p(d_omega), _p(d_a), _p(d_b), _p(d_amp),
            _p(d_bkg), G, _p(d_delays), _p(d_ll)))

    def sky_credible_dev(self, P, S, d_ll, d_log_weight, d_prob, d_level, d_order, d_log_norm):
        _check(self._lib, self._lib.ipn_sky_credible_dev(self._h, P, S, _p(d_ll), _p(d_log_weight) if d_log_weight else None,
                                                          _p(d_prob), _p(d_level), _p(d_order), _p(d_log_norm)))

    def poisson_counts_dev(self, T, d_exposure, S, d_rates, d_bkg, seed, d_out):
        _check(self._lib, self._lib.ipn_poisson_counts_dev(self._h, T, _p(d_exposure), S, _p(d_rates), _p(d_bkg),
                                                            C.c_uint64(int(seed) & (2**64 - 1)), _p(d_out)))


_default_ctx = {}
_ctx_lock = threading.Lock()


def default_context(device: int | None = None) -> Context:
    """Process-wide context per device (LOCAL_RANK selects the device under torchrun)."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    with _ctx_lock:
        ctx = _default_ctx.get(device)
        if ctx is None:
            ctx = Context(device)
            _default_ctx[device] = ctx
        return ctx
