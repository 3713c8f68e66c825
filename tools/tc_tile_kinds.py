"""Cycles per tile of the hot kernel for the two kinds of tiles: all bins with counts (log needed) and all bins empty.
`IPN_TC_DEBUG=1 python tools/tc_tile_kinds.py` (needs a B200)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from pyipn_b200 import Context
from pyipn_b200._lib import HOOKS_LIB_PATH, OPT_HIFI
from pyipn_b200.synth import delay_grid_workload
w = delay_grid_workload(T=100_000, G=4096, S=32)
ctx = Context(0, lib_path=HOOKS_LIB_PATH)  # the -DIPN_TC_HOOKS build: IPN_TC_DEBUG exists only there
ctx.set_option(OPT_HIFI, 2)
frac = float((w["counts"] > 0).mean())
for name, counts in (("as generated (%.0f %% of bins have counts)" % (100 * frac), w["counts"]), ("all bins empty", np.zeros_like(w["counts"])),
                     ("all bins with counts", np.maximum(w["counts"], 1))):
    for _ in range(2):
        ctx.loglik_delay_grid(w["time"], w["exposure"], counts, w["omega"], w["a"], w["b"], w["amp"], w["bkg"], w["delays"])
    print(name, "main kernel ms", ctx.last_main_kernel[0], flush=True)
