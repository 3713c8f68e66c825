"""Correctness probe of the tcgen05 path against the oracle (small shapes first; bounded by `timeout`)."""
import json, sys
import numpy as np
sys.path.insert(0, ".")
from oracle import ipn_oracle as orc
from pyipn_b200 import Context
from pyipn_b200._lib import OPT_LOGLIK_IMPL
from pyipn_b200.synth import delay_grid_workload

ctx = Context(0)
def run(T, G, S, impl, dt_bin=1e-3, sub=None):
    w = delay_grid_workload(T=T, G=G, S=S, dt_bin=dt_bin)
    ctx.set_option(OPT_LOGLIK_IMPL, impl)
    ll = ctx.loglik_delay_grid(w["time"], w["exposure"], w["counts"], w["omega"], w["a"], w["b"], w["amp"], w["bkg"], w["delays"])
    idx = np.arange(G) if sub is None else np.unique(np.linspace(0, G - 1, sub).astype(int))
    ns = min(S, 2)
    exp = orc.loglik_delay_grid(w["counts"], w["time"], w["exposure"], w["omega"][:ns], w["a"][:ns], w["b"][:ns], w["amp"][:ns], w["bkg"][:ns], w["delays"][idx])
    err = ll[idx][:, :ns] - exp
    print(json.dumps(dict(T=T, G=G, S=S, impl=impl, ms=ctx.last_kernel_ms, main_ms=ctx.last_main_kernel[0], err_max=float(np.abs(err).max()),
                          err_rms=float(np.sqrt((err**2).mean())), err_mean=float(err.mean()), ll0=float(ll[0, 0]), exp0=float(exp[0, 0]))), flush=True)
    return ll

run(64, 128, 1, 2, dt_bin=0.01)
run(200, 130, 2, 2, dt_bin=0.01)
run(4096, 300, 3, 2, dt_bin=0.005)
run(100_000, 4096, 4, 2, sub=48)
run(100_000, 4096, 4, 1, sub=48)
ll = run(100_000, 4096, 64, 2, sub=24)
T, G, S = 100_000, 4096, 64
print("units/s", T * G * S / (ctx.last_kernel_ms * 1e-3), "main", T * G * S / (ctx.last_main_kernel[0] * 1e-3))
