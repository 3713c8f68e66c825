"""Host mirror of ``pyipn/lightcurve.py`` (containers feeding the hot path); matplotlib/astropy-free.

``LightCurve.get_binned_light_curve`` follows pyipn/lightcurve.py:24-47 exactly (``np.arange`` edges +
``np.histogram``); ``BinnedLightCurve`` follows pyipn/lightcurve.py:96-180 (``res_ms`` without astropy units).
"""
from __future__ import annotations

import numpy as np


class LightCurve(object):
    def __init__(self, source_arrival_times, bkg_arrival_times):
        self._arrival_times = np.concatenate([source_arrival_times, bkg_arrival_times])
        self._source_arrival_time = source_arrival_times
        self._bkg_arrival_times = bkg_arrival_times

    def get_binned_light_curve(self, tstart, tstop, dt, device=False):
        """lightcurve.py:24-47.  ``device=True`` bins on the GPU (``ipn_bin_events``, bit-identical counts)."""
        if device:
            from ._lib import default_context

            return default_context().bin_events(self._arrival_times, tstart, tstop, dt)
        bins = np.arange(tstart, tstop, dt)
        counts, edges = np.histogram(self._arrival_times, bins=bins)
        rate = counts / float(dt)
        return rate, edges, counts

    @property
    def source_arrival_times(self):
        return self._source_arrival_time

    @property
    def bkg_arrival_times(self):
        return self._bkg_arrival_times


class BinnedLightCurve(object):
    def __init__(self, counts, time_bins, tstart, tstop, dt):
        assert len(counts) == len(time_bins) - 1
        assert dt > 0
        self._counts = counts
        self._time_bins = time_bins
        self._dt = dt
        self._dt_ms = dt * 1000.0
        self._tstart = tstart
        self._tstop = tstop
        self._n_bins = len(counts)

    counts = property(lambda self: self._counts)
    time_bins = property(lambda self: self._time_bins)
    dt = property(lambda self: self._dt)
    res_ms = property(lambda self: self._dt_ms)
    tstart = property(lambda self: self._tstart)
    tstop = property(lambda self: self._tstop)
    n_bins = property(lambda self: int(self._n_bins))

    @property
    def mid_points(self):
        return np.mean([self._time_bins[:-1], self._time_bins[1:]], axis=0)  # universe.py:472

    @property
    def exposure(self):
        return self._time_bins[1:] - self._time_bins[:-1]  # universe.py:473

    def time2idx(self, t):
        return np.searchsorted(self._time_bins, t)

    def get_src_counts(self, bkg_rate=500.0):
        return self._counts - bkg_rate * self._dt

    def get_max_sn(self, dt_ms, bkg_rate=500):
        """pyipn/lightcurve.py:146-173 (first maximum of the sliding-window source counts), vectorised."""
        len_int = int(dt_ms // self._dt_ms)
        cs = np.cumsum(self.get_src_counts(bkg_rate))
        n = cs.size - len_int
        if n <= 0:
            return 0, cs.size
        dif = cs[len_int:len_int + n] - cs[:n]
        i1 = int(np.argmax(dif))
        if not dif[i1] > 0.0:
            return 0, cs.size
        return i1, i1 + len_int

    @classmethod
    def from_lightcurve(cls, lightcurve, tstart, tstop, dt):
        _, t, c = lightcurve.get_binned_light_curve(tstart, tstop, dt)
        return cls(c, t, tstart, tstop, dt)
