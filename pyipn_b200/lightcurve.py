"""Event lists and binned light curves: the containers either side of the hot path (host mirror of ``pyipn/lightcurve.py``,
same class, method and property names; no matplotlib / astropy).

* ``LightCurve.get_binned_light_curve`` has the semantics of pyipn/lightcurve.py:24-47: edges ``np.arange(tstart, tstop, dt)``,
  the last bin closed on the right as ``np.histogram`` does it; returns ``(rate, edges, counts)``.  With ``device=True`` the
  events are binned by ``ipn_bin_events`` on the GPU (bit-identical counts, tests/test_gpu_configs.py).
* ``BinnedLightCurve`` carries what pyipn/lightcurve.py:96-180 carries (``res_ms`` is ``dt`` in milliseconds, formed without
  astropy units); ``mid_points`` / ``exposure`` are the per-bin arrays ``to_stan_data`` derives (pyipn/universe.py:472-473).
"""
from __future__ import annotations

import numpy as np

_MS_PER_S = 1000.0


def _host_histogram(events, tstart, tstop, dt):
    edges = np.arange(tstart, tstop, dt)
    counts, edges = np.histogram(events, bins=edges)
    return counts / float(dt), edges, counts


class LightCurve(object):
    """Arrival times of one detector: source events and background events, kept apart and merged."""

    def __init__(self, source_arrival_times, bkg_arrival_times):
        self._events = {"source": source_arrival_times, "bkg": bkg_arrival_times}
        self._arrival_times = np.concatenate((source_arrival_times, bkg_arrival_times))

    source_arrival_times = property(lambda self: self._events["source"])
    bkg_arrival_times = property(lambda self: self._events["bkg"])

    def get_binned_light_curve(self, tstart, tstop, dt, device=False):
        if not device:
            return _host_histogram(self._arrival_times, tstart, tstop, dt)
        from ._lib import default_context

        return default_context().bin_events(self._arrival_times, tstart, tstop, dt)


class BinnedLightCurve(object):
    """Counts on a grid of ``len(counts) + 1`` bin edges of width ``dt`` seconds."""

    _FIELDS = ("counts", "time_bins", "tstart", "tstop", "dt")

    def __init__(self, counts, time_bins, tstart, tstop, dt):
        assert len(time_bins) == len(counts) + 1, "one more edge than bins"
        assert dt > 0, "positive bin width"
        self._state = dict(zip(self._FIELDS, (counts, time_bins, tstart, tstop, dt)))

    @classmethod
    def from_lightcurve(cls, lightcurve, tstart, tstop, dt):
        _, edges, counts = lightcurve.get_binned_light_curve(tstart, tstop, dt)
        return cls(counts, edges, tstart, tstop, dt)

    def __getattr__(self, name):
        # counts, time_bins, tstart, tstop, dt: read-only views of the constructor's arguments
        if name in BinnedLightCurve._FIELDS:
            state = self.__dict__.get("_state")
            if state is not None:
                return state[name]
        raise AttributeError(name)

    @property
    def n_bins(self):
        return int(len(self.counts))

    @property
    def res_ms(self):
        return self.dt * _MS_PER_S

    @property
    def mid_points(self):
        lo, hi = self.time_bins[:-1], self.time_bins[1:]
        return np.mean([lo, hi], axis=0)  # the reference's expression (universe.py:472): same rounding

    @property
    def exposure(self):
        return np.diff(self.time_bins)

    def time2idx(self, t):
        """Index of the first edge that is not below ``t``."""
        return np.searchsorted(self.time_bins, t)

    def get_src_counts(self, bkg_rate=500.0):
        """Counts above a constant background of ``bkg_rate`` per second."""
        return self.counts - bkg_rate * self.dt

    def get_max_sn(self, dt_ms, bkg_rate=500):
        """Window of ``dt_ms`` milliseconds with the most background-subtracted counts, as ``(first bin, end bin)``: the FIRST
        such window if several tie, and the whole curve if no window is positive (the outcome of the scan in
        pyipn/lightcurve.py:146-173), from one prefix sum instead of a Python loop."""
        width = int(dt_ms // self.res_ms)
        prefix = np.cumsum(self.get_src_counts(bkg_rate))
        whole = (0, prefix.size)
        n_windows = prefix.size - width
        if n_windows <= 0:
            return whole
        gain = prefix[width:width + n_windows] - prefix[:n_windows]
        first = int(np.argmax(gain))
        return (first, first + width) if gain[first] > 0.0 else whole
