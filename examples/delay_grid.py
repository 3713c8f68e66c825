"""BASELINE config 2 end to end on one GPU: 2-detector delay grid, 4096 trial delays x 1e5 bins of 1 ms, K = 50, ONE draw.

  ipn_loglik_delay_grid (host buffers in, log-likelihood per delay out)  ->  argmax, gap to the second best, recovered delay

The single-draw grid is 4.1e8 (grid-point x bin) units -- under a millisecond of the contraction kernel -- so this is the
latency-bound end of the path (operand images, ~10 small launches, two host<->device copies); config 3 (bench.py) is the
throughput end.  Prints one JSON line; times are the median of the repeats (host wall clock around the call, and the
library's own CUDA-event time of the device work).

    python examples/delay_grid.py [--repeats N] [--draws S]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pyipn_b200 import Context  # noqa: E402
from pyipn_b200.likelihood import argmax_with_gap  # noqa: E402
from pyipn_b200.synth import delay_grid_workload  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--repeats", type=int, default=7)
    ap.add_argument("--draws", type=int, default=1)
    ap.add_argument("--delays", type=int, default=4096)
    ap.add_argument("--bins", type=int, default=100_000)
    a = ap.parse_args()
    w = delay_grid_workload(T=a.bins, G=a.delays, S=a.draws)
    ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    args = (w["time"], w["exposure"], w["counts"], w["omega"], w["a"], w["b"], w["amp"], w["bkg"], w["delays"])
    wall, dev, main_ms = [], [], []
    ll = None
    for i in range(a.repeats + 1):  # first call allocates the scratch buffers: not timed
        t0 = time.perf_counter()
        ll = ctx.loglik_delay_grid(*args)
        dt = time.perf_counter() - t0
        if i:
            wall.append(dt)
            dev.append(ctx.last_kernel_ms)
            main_ms.append(ctx.last_main_kernel[0])
    idx, gap = argmax_with_gap(ll)
    units = float(a.delays) * a.draws * a.bins
    t_wall, t_dev, t_main = float(np.median(wall)), float(np.median(dev)) * 1e-3, float(np.median(main_ms)) * 1e-3
    print(json.dumps({
        "workload": f"config2: {a.delays} delays x {a.bins} bins of 1 ms, K=50, {a.draws} draw(s), host buffers in/out",
        "call_s": t_wall, "device_s": t_dev, "contraction_kernel_s": t_main,
        "units_per_s_call": units / t_wall, "units_per_s_contraction": units / t_main if t_main > 0 else None,
        "argmax_index": int(idx[0]), "recovered_delay_s": float(w["delays"][idx[0]]), "true_delay_s": float(w["true_delay"]),
        "grid_step_s": float(w["delays"][1] - w["delays"][0]), "gap_to_second_best_nats": float(gap[0]),
        "loglik_impl": ctx.last_loglik_impl, "repeats": a.repeats,
    }))


if __name__ == "__main__":
    main()
