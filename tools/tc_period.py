"""Per-tile cycle budget of the tcgen05 likelihood kernel: `IPN_TC_DEBUG=1 python tools/tc_period.py S` prints the mean
cycles per tile and the barrier waits of each role (needs a B200).  IPN_TC_SKIP_MATH / IPN_TC_SKIP_MMA isolate the roles.
A second argument selects the epilogue (IPN_OPT_HIFI: 2 = plain fp32, the default here; 4 = the opt-in product epilogue)."""
import sys
sys.path.insert(0, ".")
from pyipn_b200 import Context
from pyipn_b200._lib import HOOKS_LIB_PATH, OPT_HIFI
from pyipn_b200.synth import delay_grid_workload
S = int(sys.argv[1])
w = delay_grid_workload(T=100_000, G=4096, S=S)
ctx = Context(0, lib_path=HOOKS_LIB_PATH)  # the -DIPN_TC_HOOKS build: IPN_TC_DEBUG exists only there
ctx.set_option(OPT_HIFI, int(sys.argv[2]) if len(sys.argv) > 2 else 2)
for _ in range(2):
    ll = ctx.loglik_delay_grid(w["time"], w["exposure"], w["counts"], w["omega"], w["a"], w["b"], w["amp"], w["bkg"], w["delays"])
print("ok", S, ctx.last_kernel_ms, ctx.last_main_kernel)
