"""First GPU contact: accuracy of both W modes vs the oracle on a config-2 subsample, timings, micro-peaks."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
from oracle import ipn_oracle as orc
from pyipn_b200 import Context
from pyipn_b200._lib import OPT_PRECISE_W
from pyipn_b200.synth import delay_grid_workload

ctx = Context(0)
print("fma TFLOP/s", ctx.microbench(0), "mufu Tops/s", ctx.microbench(1), flush=True)
w = delay_grid_workload(T=100_000, G=4096, S=4)
idx = np.arange(0, 4096, 64)
t0 = time.time()
exp = orc.loglik_delay_grid(w["counts"], w["time"], w["exposure"], w["omega"][:2], w["a"][:2], w["b"][:2], w["amp"][:2],
                            w["bkg"][:2], w["delays"][idx])
print("oracle s", time.time() - t0, "units/s", exp.size * 1e5 / (time.time() - t0), flush=True)
for precise in (1, 0):
    ctx.set_option(OPT_PRECISE_W, precise)
    for rep in range(2):
        t0 = time.time()
        ll = ctx.loglik_delay_grid(w["time"], w["exposure"], w["counts"], w["omega"], w["a"], w["b"], w["amp"], w["bkg"], w["delays"])
        wall = time.time() - t0
    units = ll.size * 1e5
    err = np.abs(ll[idx][:, :2] - exp)
    print(json.dumps(dict(precise=precise, kernel_ms=ctx.last_kernel_ms, wall_s=wall, units_per_s=units / (ctx.last_kernel_ms * 1e-3),
                          frac_fp32_nominal=units / (ctx.last_kernel_ms * 1e-3) * 400 / 74.4e12,
                          err_max=float(err.max()), err_rms=float(np.sqrt((err**2).mean())), err_mean=float((ll[idx][:, :2] - exp).mean()),
                          argmax=[int(i) for i in np.argmax(ll, axis=0)], launches=ctx.launch_count)), flush=True)
# larger S for a steadier timing
w = delay_grid_workload(T=100_000, G=4096, S=64)
for precise in (1, 0):
    ctx.set_option(OPT_PRECISE_W, precise)
    for rep in range(2):
        ll = ctx.loglik_delay_grid(w["time"], w["exposure"], w["counts"], w["omega"], w["a"], w["b"], w["amp"], w["bkg"], w["delays"])
    units = ll.size * 1e5
    print(json.dumps(dict(S=64, precise=precise, kernel_ms=ctx.last_kernel_ms, units_per_s=units / (ctx.last_kernel_ms * 1e-3),
                          frac_fp32_nominal=units / (ctx.last_kernel_ms * 1e-3) * 400 / 74.4e12)), flush=True)
# rff_eval timing
S, T = 64, 100_000
for rep in range(2):
    lam = ctx.rff_eval(w["time"], w["omega"], w["a"], w["b"], np.zeros(S), w["amp"])
print("rff_eval ms", ctx.last_kernel_ms, "draw*bins/s", S * T / (ctx.last_kernel_ms * 1e-3))
