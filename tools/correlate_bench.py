"""N1 throughput: the chi^2 cross-correlation scan (correlation.py:247-372) on the GPU, and -- with --reference, in a container
that has /root/reference -- the reference's own numba function on the same problem (imported with the stubs of
tests/golden/make_golden_correlation.py; the GPU box has no reference checkout, so the two halves run where they can).

    python tools/correlate_bench.py              # GPU (needs a B200): prints one JSON line
    python tools/correlate_bench.py --reference  # CPU, the reference's numba `correlate`: prints one JSON line
Problem: 20 000 trial shifts x 2 000 bins of light curve 2 (31.5 ms) against light curve 1 at 2 ms -- 6.3e8 bin additions."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def problem():
    rng = np.random.default_rng(4)
    n1, n2 = 60_000, 2_600
    dt1, dt2 = 0.002, 0.0315
    t1, t2 = np.arange(n1 + 1) * dt1, np.arange(n2 + 1) * dt2
    prof = lambda t: 300.0 * np.exp(-0.5 * ((t - 40.0) / 6.0) ** 2) * (1 + 0.4 * np.sin(t * 1.7))
    c1 = rng.poisson((prof(t1[:-1] + 3.3) + 500.0) * dt1).astype(np.float64)
    c2 = rng.poisson((prof(t2[:-1]) + 500.0) * dt2).astype(np.float64)
    s1, s2 = c1 - 500.0 * dt1, c2 - 500.0 * dt2
    return (t1, s1, c1, t2, s2, c2, 300, 2299, 100, 20_000, 1.0, dt1 * 1000.0, dt2 * 1000.0)


def main():
    args = problem()
    shifts, bins = args[9], args[7] - args[6] + 1
    work = shifts * bins
    if "--reference" in sys.argv:
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_corr")
        from make_golden_correlation import import_reference

        _, ref = import_reference()
        ref.correlate(*args[:9], 8, *args[10:])  # JIT
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            out = ref.correlate(*args)
            ts.append(time.perf_counter() - t0)
        print(json.dumps({"impl": "reference numba correlate (1 thread, as the reference runs it)", "seconds": min(ts),
                          "shift_bins_per_s": work / min(ts), "argmin": int(out[6])}))
        return
    from pyipn_b200 import Context

    ctx = Context(0)
    ts, ks = [], []
    for _ in range(5):
        t0 = time.perf_counter()
        arr_dt, arr_chi = ctx.correlate(*args)
        ts.append(time.perf_counter() - t0)
        ks.append(ctx.last_kernel_ms)
    print(json.dumps({"impl": "pyipn_b200 correlate_kernel (host buffers in and out)", "seconds": min(ts), "kernel_ms": min(ks),
                      "shift_bins_per_s": work / min(ts), "shift_bins_per_s_kernel": work / (min(ks) * 1e-3),
                      "argmin": int(np.argmin(arr_chi))}))


if __name__ == "__main__":
    main()
