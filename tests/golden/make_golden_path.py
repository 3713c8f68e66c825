"""Mint ``path_golden.npz``: outputs of the REFERENCE's own functions either side of the RFF intensity, run in the build
container (reads /root/reference; the GPU box only sees the committed .npz).

    python tests/golden/make_golden_path.py

``import pyipn`` is impossible here (astropy / arviz / ligo.skymap / h5py / ipyvolume are absent), so the functions are
lifted out of the reference files by name with ``ast`` and executed UNMODIFIED except for what cannot run here:

* ``pyipn/fit.py:651-705``  ``_expeced_rate``, ``_expeced_rate_multiscale``, ``ppc_generator`` -- compiled by numba as in
  the reference (``cache=True`` dropped so that nothing is written into /root/reference), calling the reference's own
  ``RFF`` / ``RFF_multiscale`` (pyipn/rff.py, loaded by file path).  ``ppc_generator`` is run twice: compiled (numba's
  MT19937, seeded from compiled code) and as plain Python (numpy's legacy global RandomState) -- the two streams agree.
* ``pyipn/possion_gen.py:159-242``  ``source_poisson_generator``, ``background_poisson_generator`` -- their jitclass
  container ``VectorFloat64`` no longer compiles under numba 0.65 (numba_array.py:106), so the two function bodies run
  as plain Python with a three-line list-backed stand-in for the container; ``pulse`` is the reference's compiled one.
* ``pyipn/lightcurve.py:6-47``  ``LightCurve.get_binned_light_curve`` -- the module is imported unmodified (matplotlib /
  astropy.units stubbed, as in make_golden_correlation.py).
* ``pyipn/universe.py:441-529``  ``Universe.to_stan_data`` -- the method is lifted out and bound to a stand-in object that
  only carries ``_detectors`` (location / pointing / effective-area attributes) and ``_light_curves`` (reference
  ``LightCurve`` objects): spacecraft positions are inputs, so no astropy frame arithmetic is involved.
"""
import ast
import collections
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/pyipn"
sys.path.insert(0, HERE)


def lift(path, names, drop_decorators=False, cls=None):
    """Source text of the named top-level functions (or methods of ``cls``) of a reference file."""
    src = open(path).read()
    tree = ast.parse(src)
    body = tree.body
    if cls is not None:
        body = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls).body
    out = []
    for node in body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            lines = src.splitlines()[(node.lineno if drop_decorators else min([node.lineno] + [d.lineno for d in node.decorator_list])) - 1: node.end_lineno]
            text = "\n".join(lines)
            if cls is not None:
                import textwrap

                text = textwrap.dedent(text)
            out.append(text.replace("cache=True", "cache=False"))
    assert len(out) == len(names), (path, names)
    return "\n\n".join(out)


def main():
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nbcache")
    import numba as nb

    from make_golden import load_ref, make_params
    from make_golden_correlation import import_reference

    ref_rff = load_ref()
    ref_lc, _ = import_reference()  # pyipn.lightcurve, unmodified
    rng = np.random.default_rng(20261020)
    out = {}

    # ---- fit.py draw loops and ppc_generator, compiled by numba as the reference does -----------------------------
    ns = dict(np=np, nb=nb, RFF=ref_rff.RFF, RFF_multiscale=ref_rff.RFF_multiscale)
    exec(lift(f"{REF}/fit.py", ["_expeced_rate", "_expeced_rate_multiscale", "ppc_generator"]), ns)
    k, S = 50, 9
    time = np.sort(rng.uniform(-4.0, 30.0, size=333))
    o1, o2, b1, b2, sc = make_params(rng, k, S, 34.0)
    bw = np.ones(S)  # Fit.expected_rate always passes ones (fit.py:275)
    amp = np.exp(rng.normal(0.0, 0.5, size=S))
    dt = rng.normal(0.0, 0.4, size=S)
    out["F_time"], out["F_omega1"], out["F_omega2"], out["F_beta1"], out["F_beta2"] = time, o1, o2, b1, b2
    out["F_scale"], out["F_bw"], out["F_amp"], out["F_dt"] = sc, bw, amp, dt
    out["F_rate_single"] = ns["_expeced_rate"](time, o1, o2, b1, b2, bw, sc[1], amp, dt, S)
    out["F_rate_multi"] = ns["_expeced_rate_multiscale"](time, o1, o2, b1, b2, bw, sc, amp, dt, S)

    rates = 40.0 * out["F_rate_multi"][:4, :60]
    bkg = np.array([500.0, 20.0, 3.0, 0.0])
    expo = rng.uniform(0.05, 0.3, size=60)

    @nb.njit
    def seed_numba(s):
        np.random.seed(s)

    seed_numba(4242)
    out["P_counts_numba"] = ns["ppc_generator"](time[:60], expo, rates, bkg, 4)
    np.random.seed(4242)
    out["P_counts_numpy"] = ns["ppc_generator"].py_func(time[:60], expo, rates, bkg, 4)
    out["P_rates"], out["P_bkg"], out["P_exposure"], out["P_seed"] = rates, bkg, expo, np.int64(4242)
    assert np.array_equal(out["P_counts_numba"], out["P_counts_numpy"]), "numba and numpy legacy streams differ"

    # ---- possion_gen.py thinning generators (bodies as plain Python, reference's compiled pulse) --------------------
    pkg = sys.modules["pyipn"]
    na = types.ModuleType("pyipn.numba_array")

    class VectorFloat64:  # stand-in for the jitclass container (numba_array.py): append + .arr
        def __init__(self, n):
            self._v = []

        def append(self, x):
            self._v.append(float(x))

        @property
        def arr(self):
            return np.array(self._v, dtype=np.float64)

    na.VectorFloat64 = VectorFloat64
    sys.modules["pyipn.numba_array"] = na
    pkg.numba_array = na
    import importlib

    pg = importlib.import_module("pyipn.possion_gen")  # unmodified module; nothing compiles until called
    gen_ns = dict(np=np, nb=nb, pulse=pg.pulse, VectorFloat64=VectorFloat64)
    exec(lift(f"{REF}/possion_gen.py", ["source_poisson_generator", "background_poisson_generator"], drop_decorators=True), gen_ns)
    # template_3.yaml's burst (K=500, t_rise=0.5, t_decay=4.1), seeds as Universe assigns them (seed + 10 i, + 1 for bkg)
    out["G_src"] = gen_ns["source_poisson_generator"](-2.0, 8.0, 500.0, 0.0, 0.5, 4.1, 12345)
    out["G_src_args"] = np.array([-2.0, 8.0, 500.0, 0.0, 0.5, 4.1, 12345])
    out["G_bkg"] = gen_ns["background_poisson_generator"](-2.0, 8.0, 0.0, 500.0, 12346)
    out["G_bkg_args"] = np.array([-2.0, 8.0, 0.0, 500.0, 12346])
    out["G_bkg_slope"] = gen_ns["background_poisson_generator"](0.0, 4.0, 25.0, 300.0, 77)
    out["G_bkg_slope_args"] = np.array([0.0, 4.0, 25.0, 300.0, 77])
    Ks, ts, tr, td = np.array([300.0, 150.0]), np.array([0.0, 2.5]), np.array([0.3, 0.6]), np.array([2.0, 1.5])
    out["G_src_multi"] = gen_ns["source_poisson_generator"](-1.0, 7.0, Ks, ts, tr, td, 99)
    out["G_src_multi_pulses"] = np.stack([Ks, ts, tr, td])
    out["G_src_late"] = gen_ns["source_poisson_generator"](0.0, 5.0, 500.0, 100.0, 0.5, 4.1, 1)  # pulse after the window
    x = np.linspace(-1.0, 9.0, 41)
    out["G_pulse_x"] = x
    out["G_pulse"] = pg.pulse(x, 500.0, 0.0, 0.5, 4.1)
    out["G_pulse_multi"] = pg.pulse(x, Ks, ts, tr, td)

    # ---- lightcurve.py binning and universe.py to_stan_data --------------------------------------------------------
    lc0 = ref_lc.LightCurve(out["G_src"], out["G_bkg"])
    lc1 = ref_lc.LightCurve(out["G_src_multi"], out["G_bkg_slope"])
    rate, edges, counts = lc0.get_binned_light_curve(-1.0, 6.0, 0.2)
    out["L_rate"], out["L_edges"], out["L_counts"] = rate, edges, counts
    out["L_window"] = np.array([-1.0, 6.0, 0.2])

    Det = collections.namedtuple("Det", "location pointing effective_area")

    class _Loc:
        def __init__(self, xyz):
            self._xyz = np.asarray(xyz, dtype=np.float64)

        def get_cartesian_coord(self):
            return types.SimpleNamespace(xyz=types.SimpleNamespace(value=self._xyz))

    sc_pos = np.array([[6871.0, 120.5, -33.25], [-1.2e5, 8.0e4, 3.1e4]])
    pointing = np.array([[0.0, 0.6, 0.8], [1.0, 0.0, 0.0]])
    area = np.array([1.0, 0.5])
    fake = types.SimpleNamespace(
        _detectors=collections.OrderedDict(
            (f"det{i}", Det(_Loc(sc_pos[i]), types.SimpleNamespace(cartesian=pointing[i]),
                            types.SimpleNamespace(effective_area=area[i]))) for i in range(2)),
        _light_curves={"det0": lc0, "det1": lc1},
    )
    sd_ns = dict(np=np)
    exec(lift(f"{REF}/universe.py", ["to_stan_data"], cls="Universe"), sd_ns)
    cases = {"D1": dict(tstart=-1.0, tstop=6.0, dt=0.2, k=50),  # one window for every detector
             "D2": dict(tstart=[-1.0, 0.5], tstop=[6.0, 3.75], dt=[0.2, 0.05], k=25, n_cores=2, factor=0.5, factor2=1)}  # ragged
    for tag, kw in cases.items():
        sd = sd_ns["to_stan_data"](fake, **kw)
        for key, v in sd.items():
            out[f"{tag}_{key}"] = np.asarray(v)
    out["D_sc_pos"], out["D_pointing"], out["D_area"] = sc_pos, pointing, area

    np.savez_compressed(os.path.join(HERE, "path_golden.npz"), **out)
    print("wrote path_golden.npz:", {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    sys.exit(main())
