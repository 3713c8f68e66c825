"""Host facade keeping pyipn's ``Universe`` / ``GRB`` / ``Detector`` API for the data path into the hot path.

Kept: the YAML schema and ``from_dict`` / ``from_yaml`` (pyipn/universe.py:267-342), detector registration,
wave-front arrival ordering and ``T0`` (pyipn/universe.py:109-179), light-curve creation with per-detector
seeds ``seed + 10 i`` (pyipn/universe.py:181-203, pyipn/detector.py:129-222) and ``to_stan_data``
(pyipn/universe.py:441-529) -- the data contract the likelihood kernels consume.

Also kept, astropy-free on km vectors: the accessors of ``GRB`` / ``Detector`` (grb.py:35-64, detector.py:34-127),
``calculate_annulus`` / ``localize_GRB`` (universe.py:364-386, 670-697) and ``table`` (plain dicts instead of pandas).

Not kept (out of scope): astropy frames, plotting, earth blockage, HDF5 save files.
Positions come from plain spherical trigonometry (GCRS ra/dec/altitude -> km); astropy's ICRS->GCRS
aberration terms are NOT applied, which changes T0 at the 1e-3 relative level.
"""
from __future__ import annotations

import collections

import numpy as np

from .lightcurve import LightCurve
from .possion_gen import background_poisson_generator, source_poisson_generator

C_KM_S_ASTROPY = 299792.458  # astropy.constants.c, used by the reference for T0 (universe.py:146-149)
EARTH_RADIUS_KM = 6371.0  # geometry.py:88
MPC_KM = 3.0856775814913673e19  # 1 Mpc in km (geometry.py:80-84 builds the GRB position in Mpc)


def _unit_vector(ra_deg, dec_deg):
    ra, dec = np.deg2rad(ra_deg), np.deg2rad(dec_deg)
    return np.array([np.cos(dec) * np.cos(ra), np.cos(dec) * np.sin(ra), np.sin(dec)])


class GRB(object):
    """pyipn/grb.py:8-64: location + Norris pulse parameters (scalars, or arrays for multi-pulse)."""

    def __init__(self, ra, dec, distance, K, t_rise, t_decay, t_start=None):
        self.ra, self.dec, self.distance = ra, dec, distance
        self._K, self._t_rise, self._t_decay, self._t_start = K, t_rise, t_decay, t_start

    @property
    def pulse_parameters(self):
        return self._K, self._t_rise, self._t_decay, self._t_start

    @property
    def norm_vec(self):
        return _unit_vector(self.ra, self.dec)

    @property
    def location(self):
        """grb.py:46: here the Cartesian position in km (the reference returns an astropy-backed GRBLocation)."""
        return self.distance * MPC_KM * self.norm_vec

    @property
    def table(self):
        """grb.py:50-64 as a plain dict (pandas is not a dependency of the path)."""
        K, t_rise, t_decay, t_start = self.pulse_parameters
        return collections.OrderedDict([("ra", self.ra), ("dec", self.dec), ("distance", self.distance), ("K", K),
                                        ("t_rise", t_rise), ("t_decay", t_decay), ("t_start", t_start)])


class Detector(object):
    """pyipn/detector.py:10-222 reduced to what the data path needs."""

    def __init__(self, ra, dec, altitude, pointing_ra, pointing_dec, effective_area, name):
        self.name = name
        self.ra, self.dec, self.altitude = ra, dec, altitude
        self.pointing_ra, self.pointing_dec = pointing_ra, pointing_dec
        self.total_area = effective_area
        self._background_slope = 0.0
        self._background_norm = 500.0  # detector.py:11: flat 500 cps

    @property
    def xyz_km(self):
        return (EARTH_RADIUS_KM + self.altitude) * _unit_vector(self.ra, self.dec)  # geometry.py:106-113

    @property
    def pointing_vec(self):
        return _unit_vector(self.pointing_ra, self.pointing_dec)

    # detector.py:34-78: the reference's accessors, astropy-free (vectors in km / unit vectors instead of SkyCoord)
    location = property(lambda self: self.xyz_km)
    pointing = property(lambda self: self.pointing_vec)
    effective_area = property(lambda self: self.total_area)

    def angular_separation(self, grb):
        """detector.py:98-127 / geometry.py:24-36: angle between the pointing and the GRB direction, radians."""
        return float(np.arccos(np.clip(np.dot(self.pointing_vec, grb.norm_vec), -1.0, 1.0)))

    def light_travel_time(self, grb):
        """detector.py:80-96: |detector - GRB| / c in seconds (c = astropy's 299792.458 km/s)."""
        return float(np.linalg.norm(self.xyz_km - grb.location) / C_KM_S_ASTROPY)

    def effective_area_towards(self, grb):
        """effective_area.py:49-64: A cos(theta), zero beyond 90 degrees."""
        cosang = float(np.clip(np.dot(self.pointing_vec, grb.norm_vec), -1.0, 1.0))
        return self.total_area * cosang if cosang > 0.0 else 0.0

    def build_light_curve(self, grb, T0, tstart, tstop, seed=1234):
        """detector.py:129-222 (no earth blockage)."""
        K, t_rise, t_decay, t_start = grb.pulse_parameters
        area = self.effective_area_towards(grb)
        if np.ndim(K) == 0:
            src = source_poisson_generator(tstart, tstop, K * area, T0, t_rise, t_decay, seed)
        else:
            src = source_poisson_generator(tstart, tstop, np.asarray(K) * area, np.asarray(t_start) + T0,
                                           np.asarray(t_rise), np.asarray(t_decay), seed)
        bkg = background_poisson_generator(tstart, tstop, self._background_slope, self._background_norm, seed + 1)
        return LightCurve(src, bkg)


class Universe(object):
    def __init__(self, grb, yaml_dict=None, locked=False, seed=1234):
        self._seed = seed
        self._detectors = collections.OrderedDict()
        self._grb = grb
        self._T0 = None
        self._time_differences = None
        self._light_curves = None
        self._n_detectors = 0
        self._locked = locked
        self._yaml_dict = yaml_dict

    grb = property(lambda self: self._grb)
    T0 = property(lambda self: self._T0)
    detectors = property(lambda self: self._detectors)
    light_curves = property(lambda self: self._light_curves)

    @property
    def grb_radius(self):
        return self._grb.distance  # universe.py:87-89 (Mpc)

    def calculate_annulus(self, detector1, detector2):
        """universe.py:364-386 for two detector names: (unit connection vector, its (ra, dec) in degrees, half-opening
        angle in radians) from the pair's T0 difference; plain km vectors instead of astropy frames."""
        from .sky import calculate_annulus

        names = list(self._detectors.keys())
        sc_pos = np.array([d.xyz_km for d in self._detectors.values()])
        norm_d, radec, theta = calculate_annulus(sc_pos, self._T0, names.index(detector1), names.index(detector2))
        return norm_d, np.degrees(radec), theta

    def localize_GRB(self):
        """universe.py:670-697: least-squares intersection of all pairwise annuli; returns the solution vector."""
        from .sky import localize_GRB

        return localize_GRB(np.array([d.xyz_km for d in self._detectors.values()]), self._T0)

    @property
    def table(self):
        """universe.py:699-734 as a plain dict of columns (pandas is not a dependency of the path)."""
        out = collections.OrderedDict()
        out["name"] = list(self._detectors.keys())
        out["dt"] = None if self._T0 is None else np.asarray(self._T0)
        out["altitude"] = [d.altitude for d in self._detectors.values()]
        out["position"] = [f"{d.ra},{d.dec}" for d in self._detectors.values()]
        out["pointing"] = [f"{d.pointing_ra},{d.pointing_dec}" for d in self._detectors.values()]
        out["eff. area"] = [d.total_area for d in self._detectors.values()]
        return out

    def register_detector(self, detector):
        self._detectors[detector.name] = detector
        self._n_detectors += 1
        assert self._n_detectors == len(self._detectors.keys())

    def explode_grb(self, tstart, tstop, verbose=True):
        if self._locked:
            print("This universe is locked")
            return
        self._compute_time_differences()
        self._create_light_curves(tstart, tstop)

    def _compute_time_differences(self):
        """universe.py:109-179: T0_d = cumulative light-travel lag of the wave front, detectors re-ordered by T0."""
        g = self._grb.norm_vec
        names = list(self._detectors.keys())
        ltd = np.array([-np.dot(g, self._detectors[n].xyz_km) for n in names])
        order = np.argsort(ltd, kind="stable")
        lts = ltd[order]
        dts = np.concatenate([[0.0], np.diff(lts) / C_KM_S_ASTROPY])
        assert np.all(dts >= 0)
        self._T0 = np.cumsum(dts)
        self._time_differences = dts
        new = collections.OrderedDict()
        for i in order:
            new[names[i]] = self._detectors[names[i]]
        self._detectors = new

    def _create_light_curves(self, tstart, tstop):
        self._light_curves = collections.OrderedDict()
        i = 0
        for t0, (name, det) in zip(self._T0, self._detectors.items()):
            self._light_curves[name] = det.build_light_curve(self._grb, t0, tstart, tstop, seed=self._seed + i)
            i += 10

    def set_light_curves(self, light_curves):
        """Attach externally produced LightCurve objects (ordered like ``detectors``)."""
        self._light_curves = collections.OrderedDict(light_curves)

    @classmethod
    def from_dict(cls, d, locked=False):
        gp = d["grb"]
        seed = d["seed"] if "seed" in d else 12345
        t_start = gp["t_start"] if "t_start" in gp else None
        K, t_rise, t_decay = gp["K"], gp["t_rise"], gp["t_decay"]
        if isinstance(K, list):
            K, t_rise, t_decay, t_start = (np.array(x) for x in (K, t_rise, t_decay, t_start))
        grb = GRB(gp["ra"], gp["dec"], gp["distance"], K, t_rise, t_decay, t_start)
        universe = cls(grb, yaml_dict=d, locked=locked, seed=seed)
        for name, v in d["detectors"].items():
            universe.register_detector(Detector(v["ra"], v["dec"], v["altitude"], v["pointing"]["ra"],
                                                v["pointing"]["dec"], v["effective_area"], name))
        return universe

    @classmethod
    def from_yaml(cls, yaml_file):
        import yaml

        with open(yaml_file, "r") as f:
            return cls.from_dict(yaml.load(f, Loader=yaml.SafeLoader))

    def to_stan_data(self, tstart, tstop, dt=0.2, k=50, n_cores=1, factor=0.5, factor2=1):
        """universe.py:441-529: padded (D, max_bins) int counts, bin mid-points, exposures, sc_pos (km), grainsizes."""
        n_dets = len(self._detectors)
        counts, times, exposures, n_time_bins = [], [], [], []
        sc_pos = np.empty((n_dets, 3))
        sc_pointing = np.empty((n_dets, 3))
        effective_area = np.empty(n_dets)
        tstart, tstop, dt = np.atleast_1d(tstart), np.atleast_1d(tstop), np.atleast_1d(dt)
        for n, (name, det) in enumerate(self._detectors.items()):
            lc = self._light_curves[name]
            # universe.py:466-468 falls back to the first time selection when fewer selections than detectors are
            # given.  (The reference reuses the loop variable for that, which also redirects its sc_pos row to 0 and
            # leaves the other rows uninitialised; that slip is not reproduced here.)
            m = n if n < len(tstart) else 0
            _, t, c = lc.get_binned_light_curve(tstart[m], tstop[m], dt[m])
            counts.append(c)
            times.append(np.mean([t[:-1], t[1:]], axis=0))
            exposures.append(t[1:] - t[:-1])
            n_time_bins.append(len(c))
            sc_pos[n] = det.xyz_km
            sc_pointing[n] = det.pointing_vec
            effective_area[n] = det.effective_area_towards(self._grb)
        mx = max(n_time_bins)
        counts_stan = np.zeros((n_dets, mx), dtype=int)
        times_stan = np.zeros((n_dets, mx))
        exposure_stan = np.zeros((n_dets, mx))
        for n in range(n_dets):
            counts_stan[n, : n_time_bins[n]] = counts[n]
            times_stan[n, : n_time_bins[n]] = times[n]
            exposure_stan[n, : n_time_bins[n]] = exposures[n]
        grainsize = [int(np.round(n / n_cores) * factor) for n in n_time_bins]
        return dict(
            N_detectors=n_dets, N_time_bins=n_time_bins, max_N_time_bins=mx, counts=counts_stan, time=times_stan,
            exposure=exposure_stan, effective_area=effective_area, sc_pos=sc_pos, sc_pointing_norm=sc_pointing, k=k,
            grainsize=grainsize, grainsize_meta=max(1, int(np.round(n_dets / n_cores) * factor2)),
        )
