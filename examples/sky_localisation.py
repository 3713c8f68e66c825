"""BASELINE config 4 end to end on one GPU: 8 detectors x 1e5 bins each, HEALPix nside = 64 (49 152 directions), one draw.

  ipn_loglik_sky_grid  ->  per-pixel log-likelihood (detector sum, delays from the directions)
  ipn_sky_credible     ->  probabilities, credible levels, 68 / 95 % regions
  sky.localize_GRB     ->  classical least-squares triangulation from the true delays, for comparison

Prints one JSON line (timings are host wall-clock around the library calls, second call of each).

    python examples/sky_localisation.py [--draws S]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pyipn_b200 import Context, sky  # noqa: E402
from pyipn_b200.synth import sky_grid_workload  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--draws", type=int, default=1)
    ap.add_argument("--nside", type=int, default=64)
    ap.add_argument("--bins", type=int, default=100_000)
    a = ap.parse_args()
    w = sky_grid_workload(D=8, T=a.bins, S=a.draws, nside=a.nside)
    P = w["ghat"].shape[0]
    ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    args = (w["nbins"], w["time"], w["exposure"], w["counts"], w["sc_pos"], w["ghat"], w["omega"], w["a"], w["b"], w["amp"], w["bkg"])
    for _ in range(2):
        t0 = time.perf_counter()
        ll = ctx.loglik_sky_grid(*args)
        t_grid = time.perf_counter() - t0
    for _ in range(2):
        t0 = time.perf_counter()
        out = sky.credible_regions(ll, levels=(0.68, 0.95), pixel_area=4 * np.pi / P, ctx=ctx)
        t_cred = time.perf_counter() - t0
    best = w["ghat"][out["best"]]
    tri = sky.localize_GRB(w["sc_pos"], -(w["sc_pos"] @ w["g_true"]) / sky.SPEED_OF_LIGHT_KM_S)
    tri /= np.linalg.norm(tri)
    deg = 180.0 / np.pi
    units = float(P) * 8 * a.bins * a.draws  # detector 0 is evaluated too (delay 0): 8 shifted light curves per direction
    print(json.dumps({
        "workload": f"config4: 8 detectors x {a.bins} bins, nside {a.nside} ({P} directions), {a.draws} draw(s)",
        "sky_grid_s": t_grid, "units_per_s": units / t_grid, "credible_s": t_cred,
        "best_pixel_offset_deg": float(np.arccos(np.clip(best @ w["g_true"], -1, 1)) * deg),
        "pixel_size_deg": float(np.sqrt(4 * np.pi / P) * deg),
        "area68_deg2": out["areas"][0.68] * deg * deg, "area95_deg2": out["areas"][0.95] * deg * deg,
        "pixels68": int(out["masks"][0.68].sum()), "pixels95": int(out["masks"][0.95].sum()),
        "triangulation_offset_deg": float(np.arccos(np.clip(tri @ w["g_true"], -1, 1)) * deg),
        "loglik_impl": ctx.last_loglik_impl,
    }))


if __name__ == "__main__":
    main()
