"""Short single-GPU run for ncu: ipn_rff_eval at 64 draws x 1e5 bins (J = 100)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from pyipn_b200 import Context
from pyipn_b200.synth import delay_grid_workload
w = delay_grid_workload(T=100_000, G=16, S=64)
ctx = Context(0)
for _ in range(3):
    lam = ctx.rff_eval(w["time"], w["omega"], w["a"], w["b"], np.zeros(64), w["amp"])
print("ok", lam.shape, ctx.last_kernel_ms, "ms", 64 * 1e5 / (ctx.last_kernel_ms * 1e-3), "draw*bins/s")
