#!/bin/bash
# Round-2 session 4: where does the 12-product tile spend its time?  (hooks builds: role isolation; level-6 variants)
mkdir -p gpurun_out/roles; O=gpurun_out/roles
cat > /tmp/period.py <<'PY'
import sys, os
sys.path.insert(0, ".")
from pyipn_b200 import Context
from pyipn_b200._lib import OPT_HIFI
from pyipn_b200.synth import delay_grid_workload
w = delay_grid_workload(T=100_000, G=4096, S=64)
ctx = Context(0, lib_path=os.environ["HOOKLIB"])
ctx.set_option(OPT_HIFI, 2)
for _ in range(2):
    ll = ctx.loglik_delay_grid(w["time"], w["exposure"], w["counts"], w["omega"], w["a"], w["b"], w["amp"], w["bkg"], w["delays"])
print("ok", ctx.last_kernel_ms, ctx.last_main_kernel)
PY
for v in default; do
  lib=pyipn_b200/libipn_b200_${v}_hooks.so; [ $v = default ] && lib=pyipn_b200/libipn_b200_hooks.so
  for hook in none IPN_TC_SKIP_MATH=1 IPN_TC_SKIP_MMA=1; do
    echo "== $v $hook"
    if [ $hook = none ]; then HOOKLIB=$PWD/$lib IPN_TC_DEBUG=1 timeout 120 python /tmp/period.py 2>&1 | tail -2
    else env $hook HOOKLIB=$PWD/$lib IPN_TC_DEBUG=1 timeout 120 python /tmp/period.py 2>&1 | tail -2; fi
  done
done > $O/periods.txt 2>&1
cat $O/periods.txt
