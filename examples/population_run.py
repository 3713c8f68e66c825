#!/usr/bin/env python
"""BASELINE.json config 5: population run -- N synthetic GRBs x D detectors, every stage on the GPU:
RFF intensity per detector (ipn_rff_eval) -> binned Poisson counts (ipn_poisson_counts) -> delay-grid likelihood of
every detector against the reference detector (ipn_loglik_delay_grid).  GRBs are independent: under torchrun they are
sharded over the ranks (125 per GPU for 1000 GRBs on 8 GPUs) and only the recovered delays are gathered.

    python examples/population_run.py --grbs 40            # one GPU
    torchrun --nproc-per-node 8 examples/population_run.py --grbs 1000
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pyipn_b200 import Context  # noqa: E402
from pyipn_b200.distributed import shard_bounds  # noqa: E402
from pyipn_b200.params import fold_draws  # noqa: E402
from pyipn_b200.synth import draw_prior  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grbs", type=int, default=40)
    ap.add_argument("--detectors", type=int, default=8)
    ap.add_argument("--bins", type=int, default=100_000)
    ap.add_argument("--delays", type=int, default=4096)
    ap.add_argument("--k", type=int, default=50)
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    lo, hi = shard_bounds(a.grbs, rank, world)
    ctx = Context(local)
    T, G, D = a.bins, a.delays, a.detectors
    mid = 1e-3 * (np.arange(T) + 0.5)
    expo = np.full(T, 1e-3)
    delays = np.linspace(-0.5, 0.5, G)
    err, t0 = [], time.perf_counter()
    kernel_ms = 0.0
    for grb in range(lo, hi):
        rng = np.random.default_rng(1000 + grb)  # per-GRB stream: results do not depend on the sharding
        om, b1, b2, sc = draw_prior(rng, a.k, mid[-1] - mid[0])
        w, aa, bb = fold_draws(om[0], om[1], b1, b2, 1.0, sc[:, None])
        true = np.concatenate([[0.0], rng.uniform(-0.4, 0.4, D - 1)])
        amp = 400.0 * np.exp(rng.normal(0, 0.2, D))
        rates = ctx.rff_eval(mid, np.repeat(w, D, 0), np.repeat(aa, D, 0), np.repeat(bb, D, 0), true, amp)
        kernel_ms += ctx.last_kernel_ms
        counts = ctx.poisson_counts(expo, rates, np.full(D, 500.0), seed=grb)
        kernel_ms += ctx.last_kernel_ms
        for d in range(1, D):
            ll = ctx.loglik_delay_grid(mid, expo, counts[d], w, aa, bb, amp[d:d + 1], [500.0], delays)
            kernel_ms += ctx.last_kernel_ms
            err.append(delays[int(np.argmax(ll[:, 0]))] - true[d])
    wall = time.perf_counter() - t0
    err = np.array(err)
    units = (hi - lo) * (D - 1) * G * T
    out = dict(rank=rank, world=world, grbs=hi - lo, detectors=D, wall_s=wall, kernel_s=kernel_ms * 1e-3,
               grid_units_per_s_wall=units / wall, grid_units_per_s_kernels=units / (kernel_ms * 1e-3),
               delay_error_rms_s=float(np.sqrt(np.mean(err**2))) if err.size else None,
               delay_error_max_s=float(np.abs(err).max()) if err.size else None, grid_step_s=float(delays[1] - delays[0]))
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
