"""Short single-GPU run for ncu: one delay-grid call at the bench shape with a handful of draws."""
import sys
sys.path.insert(0, ".")
from pyipn_b200 import Context
from pyipn_b200._lib import OPT_LOGLIK_IMPL
from pyipn_b200.synth import delay_grid_workload

impl = int(sys.argv[1]) if len(sys.argv) > 1 else 0
S = int(sys.argv[2]) if len(sys.argv) > 2 else 16
w = delay_grid_workload(T=100_000, G=4096, S=S)
ctx = Context(0)
ctx.set_option(OPT_LOGLIK_IMPL, impl)
for _ in range(2):
    ll = ctx.loglik_delay_grid(w["time"], w["exposure"], w["counts"], w["omega"], w["a"], w["b"], w["amp"], w["bkg"], w["delays"])
print("ok", ll.shape, float(ll.max()), ctx.last_kernel_ms, ctx.last_main_kernel)
