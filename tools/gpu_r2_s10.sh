#!/bin/bash
mkdir -p gpurun_out/r2s10; O=gpurun_out/r2s10
cat > /tmp/period.py <<'PY'
import sys, os
sys.path.insert(0, ".")
from pyipn_b200 import Context
from pyipn_b200._lib import OPT_HIFI, HOOKS_LIB_PATH
from pyipn_b200.synth import delay_grid_workload
w = delay_grid_workload(T=100_000, G=4096, S=64)
ctx = Context(0, lib_path=HOOKS_LIB_PATH)
ctx.set_option(OPT_HIFI, 2)
for _ in range(2):
    ll = ctx.loglik_delay_grid(w["time"], w["exposure"], w["counts"], w["omega"], w["a"], w["b"], w["amp"], w["bkg"], w["delays"])
print("ok", ctx.last_kernel_ms, ctx.last_main_kernel)
PY
for ct in 64 128 256 512; do echo "== chunk tiles $ct"; IPN_TC_CHUNK_TILES=$ct IPN_TC_DEBUG=1 timeout 120 python /tmp/period.py 2>&1 | tail -2; done > $O/chunks.txt 2>&1
cat $O/chunks.txt
